// offline study of orderings inside an inverted list (M = 8): wavefronts per warp-wide lookup.
// Inputs: codes.bin / off.bin / freq.bin written to /tmp/sim by scratch/make_cache.py (the bench index); build: gcc -O2 -o sim sim_reorder.c
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#define M 8
static uint8_t *codes; static int64_t off[1025], freq[1024];
// cost of one group of n<=32 entries idx[]: sum_j max_bank distinct
static int group_cost(const uint8_t *c, const int *idx, int n) {
    int tot = 0;
    for (int j = 0; j < M; ++j) {
        uint8_t seen[256]; memset(seen, 0, 256); int load[32]; memset(load, 0, sizeof load); int mx = 0;
        for (int i = 0; i < n; ++i) { int v = c[idx[i] * M + j]; if (!seen[v]) { seen[v] = 1; if (++load[v & 31] > mx) mx = load[v & 31]; } }
        tot += mx;
    }
    return tot;
}
// layout: position pos in list -> group: block b = pos/64, h = pos&1, lane = (pos%64)/2
static double order_cost(const uint8_t *c, const int *ord, int len, long *ngroups) {
    long tot = 0, ng = 0; int idx[32];
    for (int b = 0; b < len; b += 64) for (int h = 0; h < 2; ++h) {
        int n = 0; for (int p = b + h; p < len && p < b + 64; p += 2) idx[n++] = ord[p];
        if (n) { tot += group_cost(c, idx, n); ng++; }
    }
    *ngroups += ng; return (double)tot;
}
// current kernel: greedy within chunks of RCH; cost = sum_j (dup?0:2*load+1)
static void greedy(const uint8_t *c, int len, int RCH, int mode, int *ord) {
    char *alive = malloc(len);
    for (int cb = 0; cb < len; cb += RCH) {
        int n = len - cb < RCH ? len - cb : RCH; memset(alive, 1, n);
        for (int bb = 0; bb < n; bb += 64) { int rblk = n - bb < 64 ? n - bb : 64;
            for (int h = 0; h < 2; ++h) { int gsize = h == 0 ? (rblk + 1) / 2 : rblk / 2;
                int load[M][32]; uint8_t seen[M][256]; memset(load, 0, sizeof load); memset(seen, 0, sizeof seen); int mx[M]; memset(mx,0,sizeof mx);
                for (int t = 0; t < gsize; ++t) {
                    long best = -1; int bi = -1;
                    for (int i = 0; i < n; ++i) if (alive[i]) {
                        long cost = 0;
                        for (int j = 0; j < M; ++j) { int v = c[(cb + i) * M + j]; if (!seen[j][v]) { int l = load[j][v & 31];
                            if (mode == 0) cost += 2 * l + 1; else if (mode == 1) cost += (l + 1 > mx[j] ? 1000 : 0) + 2 * l + 1; else cost += (l+1>mx[j]?64:0) + l*l*l; } }
                        if (bi < 0 || cost < best) { best = cost; bi = i; }
                    }
                    alive[bi] = 0; ord[cb + bb + 2 * t + h] = cb + bi;
                    for (int j = 0; j < M; ++j) { int v = c[(cb + bi) * M + j]; if (!seen[j][v]) { seen[j][v] = 1; if (++load[j][v & 31] > mx[j]) mx[j] = load[j][v&31]; } }
                }
            } }
    }
    free(alive);
}
// balanced assignment: candidates in order of decreasing code popularity -> best group with a free slot
static void balanced(const uint8_t *c, int len, int *ord, int pen) {
    int G = 0; int gb[4096], gh[4096], gcap[4096];
    for (int b = 0; b < len; b += 64) { int rblk = len - b < 64 ? len - b : 64; for (int h = 0; h < 2; ++h) { int gs = h == 0 ? (rblk + 1) / 2 : rblk / 2; if (gs) { gb[G] = b; gh[G] = h; gcap[G] = gs; G++; } } }
    int (*load)[M][32] = calloc(G, sizeof *load); uint8_t (*seen)[M][256] = calloc(G, sizeof *seen); int (*mx)[M] = calloc(G, sizeof *mx); int *fill = calloc(G, 4);
    // popularity score
    long hist[M][256]; memset(hist, 0, sizeof hist); for (int i = 0; i < len; ++i) for (int j = 0; j < M; ++j) hist[j][c[i*M+j]]++;
    long *score = malloc(len * 8); int *perm = malloc(len * 4);
    for (int i = 0; i < len; ++i) { long sc = 0; for (int j = 0; j < M; ++j) { long bsum = 0; for (int v = c[i*M+j] & 31; v < 256; v += 32) bsum += hist[j][v]; sc += bsum; } score[i] = sc; perm[i] = i; }
    // sort perm by score descending (simple insertion-free: qsort with global)
    for (int i = 1; i < len; ++i) { int p = perm[i]; long sp = score[p]; int k = i - 1; while (k >= 0 && score[perm[k]] < sp) { perm[k+1] = perm[k]; k--; } perm[k+1] = p; }
    for (int q = 0; q < len; ++q) { int i = perm[q]; long best = -1; int bg = -1;
        for (int g = 0; g < G; ++g) if (fill[g] < gcap[g]) { long cost = 0;
            for (int j = 0; j < M; ++j) { int v = c[i*M+j]; if (!seen[g][j][v]) { int l = load[g][j][v&31]; cost += (l + 1 > mx[g][j] ? pen : 0) + 2*l + 1; } }
            if (bg < 0 || cost < best) { best = cost; bg = g; } }
        ord[gb[bg] + 2 * fill[bg] + gh[bg]] = i; fill[bg]++;
        for (int j = 0; j < M; ++j) { int v = c[i*M+j]; if (!seen[bg][j][v]) { seen[bg][j][v] = 1; if (++load[bg][j][v&31] > mx[bg][j]) mx[bg][j] = load[bg][j][v&31]; } } }
    free(load); free(seen); free(mx); free(fill); free(score); free(perm);
}
// local search: swap entries between two random groups if total cost drops
static void refine(const uint8_t *c, int len, int *ord, int iters) {
    int ngr = 0; for (int b = 0; b < len; b += 64) ngr += (len - b > 1) ? 2 : 1;
    if (ngr < 2) return;
    for (int it = 0; it < iters; ++it) {
        int p1 = rand() % len, p2 = rand() % len;
        int g1 = (p1 / 64) * 2 + (p1 & 1), g2 = (p2 / 64) * 2 + (p2 & 1); if (g1 == g2) continue;
        int i1[32], i2[32], n1 = 0, n2 = 0, k1 = -1, k2 = -1;
        for (int p = (p1/64)*64 + (p1&1); p < len && p < (p1/64)*64 + 64; p += 2) { if (p == p1) k1 = n1; i1[n1++] = ord[p]; }
        for (int p = (p2/64)*64 + (p2&1); p < len && p < (p2/64)*64 + 64; p += 2) { if (p == p2) k2 = n2; i2[n2++] = ord[p]; }
        int before = group_cost(c, i1, n1) + group_cost(c, i2, n2);
        int t = i1[k1]; i1[k1] = i2[k2]; i2[k2] = t;
        int after = group_cost(c, i1, n1) + group_cost(c, i2, n2);
        if (after < before) { t = ord[p1]; ord[p1] = ord[p2]; ord[p2] = t; }
    }
}
int main(int argc, char **argv) {
    FILE *f = fopen("off.bin", "rb"); fread(off, 8, 1025, f); fclose(f);
    f = fopen("freq.bin", "rb"); fread(freq, 8, 1024, f); fclose(f);
    codes = malloc(off[1024] * M); f = fopen("codes.bin", "rb"); fread(codes, 1, off[1024] * M, f); fclose(f);
    int nl = argc > 1 ? atoi(argv[1]) : 40;
    // sample lists proportional to freq*len (work share)
    srand(1);
    double tw = 0; for (int l = 0; l < 1024; ++l) tw += (double)freq[l] * (off[l + 1] - off[l]);
    double acc[8] = {0}; long ng[8] = {0};
    const char *names[8] = {"whole maxaware", "whole maxaware+refine", "4096 maxaware", "2048 maxaware", "balanced pen1000", "balanced pen8", "balanced pen8+refine", "1024 maxaware"};
    for (int s = 0; s < nl; ++s) {
        double r = (double)rand() / RAND_MAX * tw; int l = 0; for (; l < 1023; ++l) { r -= (double)freq[l] * (off[l + 1] - off[l]); if (r <= 0) break; }
        int len = off[l + 1] - off[l]; const uint8_t *c = codes + off[l] * M; int *ord = malloc(len * 4);
        greedy(c, len, 1 << 30, 1, ord); acc[0] += order_cost(c, ord, len, &ng[0]);
        refine(c, len, ord, len * 300); acc[1] += order_cost(c, ord, len, &ng[1]);
        greedy(c, len, 4096, 1, ord); acc[2] += order_cost(c, ord, len, &ng[2]);
        greedy(c, len, 2048, 1, ord); acc[3] += order_cost(c, ord, len, &ng[3]);
        balanced(c, len, ord, 1000); acc[4] += order_cost(c, ord, len, &ng[4]);
        balanced(c, len, ord, 8); acc[5] += order_cost(c, ord, len, &ng[5]);
        refine(c, len, ord, len * 300); acc[6] += order_cost(c, ord, len, &ng[6]);
        greedy(c, len, 1024, 1, ord); acc[7] += order_cost(c, ord, len, &ng[7]);
        free(ord);
    }
    for (int k = 0; k < 8; ++k) printf("%-24s wavefronts/lookup %.3f\n", names[k], acc[k] / (ng[k] * M));
    return 0;
}
