// Build + error-path check of the C++ host mirror (include/mmidx.hpp) -- runs WITHOUT a GPU:
//   * every class of the mirror instantiates and links against libmmidx.so;
//   * argument errors that the reference raises before touching the index are raised here too;
//   * without an sm_100 device the constructors fail loudly ("no CPU path"), never silently;
//   * RandomPermutation reproduces java.util.Random / Collections.shuffle (printed for the Python side to compare).
// With a device present (argv[1] == "gpu") it runs a tiny Linear / IVFPQ / VLAD round trip instead.
#include <cstdio>
#include <cstring>
#include <string>

#include "mmidx.hpp"

using namespace mmidx;

static int fails = 0;
#define EXPECT(cond)                                              \
    do {                                                          \
        if (!(cond)) {                                            \
            std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); \
            ++fails;                                              \
        }                                                         \
    } while (0)

template <typename F>
static std::string message_of(F f) {
    try {
        f();
    } catch (const Exception &e) {
        return e.what();
    }
    return "";
}

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
    const std::vector<int32_t> perm = random_permutation(1, 16);
    std::printf("perm");
    for (int32_t v : perm) std::printf(" %d", v);
    std::printf("\n");
    // PQ.java:148-150 is raised before the index is created
    EXPECT(message_of([] { PQ pq(10, 100, 3, 16); }) == "The given number of subvectors is not valid!");
    if (!gpu) {
        const std::string m1 = message_of([] { Linear lin(8, 100); });
        std::printf("create without device: %s\n", m1.c_str());
        EXPECT(!m1.empty());  // fails loudly: there is no CPU path
        EXPECT(!message_of([] { IVFPQ ix(16, 100, 4, 16, TransformationType::None, 8); }).empty());
        const std::vector<double> cb(4 * 3, 0.5);
        VladAggregator agg(cb, 4, 3);
        EXPECT(agg.getVectorLength() == 12);
        EXPECT(message_of([&] { agg.aggregate({{1.0, 2.0}}); }) == "Descriptor length does not match codebook centroid length");
        EXPECT(!message_of([&] { agg.aggregate({{1.0, 2.0, 3.0}}); }).empty());  // needs the device
    } else {
        Linear lin(4, 3);
        EXPECT(lin.indexVector("a", {0, 0, 0, 0}));
        EXPECT(lin.indexVector("b", {1, 0, 0, 0}));
        EXPECT(!lin.indexVector("a", {2, 0, 0, 0}));  // duplicate id -> false
        EXPECT(lin.indexVector("c", {3, 0, 0, 0}));
        EXPECT(!lin.indexVector("d", {4, 0, 0, 0}));  // full -> false
        EXPECT(message_of([&] { lin.indexVector("e", {1, 2}); }).empty());  // full is checked first (ASS.java:232)
        Answer ans = lin.computeNearestNeighbors(2, std::vector<double>{0.9, 0, 0, 0});
        EXPECT(ans.getIds().size() == 2 && ans.getIds()[0] == "b" && ans.getIds()[1] == "a");
        EXPECT(ans.getDistances()[0] == (1 - 0.9) * (1 - 0.9));
        EXPECT(lin.computeNearestNeighbors(1, std::string("c")).getIds()[0] == "c");
        EXPECT(message_of([&] { lin.computeNearestNeighbors(1, std::vector<double>{1, 2}); }) == "The dimensionality of the vector is wrong!");
        {  // RandomRotation with a supplied matrix == no transformation on the rotated vectors (RandomRotation.java:44-49)
            const std::vector<double> pqcb = {0, 0, 4, 4, 0, 0, 4, 4};  // [m=2][ks=2][2]
            const std::vector<double> R = {0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0, -1, 0, 1, 0, 0};  // signed permutation, v R
            auto rot = [&](const std::vector<double> &v) {
                std::vector<double> o(4, 0.0);
                for (int j = 0; j < 4; ++j)
                    for (int i = 0; i < 4; ++i) o[j] += v[i] * R[i * 4 + j];
                return o;
            };
            PQ a(4, 10, 2, 2, TransformationType::RandomRotation), b(4, 10, 2, 2);
            a.loadProductQuantizer(pqcb);
            b.loadProductQuantizer(pqcb);
            EXPECT(a.rotationPending());
            EXPECT(!message_of([&] { a.indexVector("x", {1, 2, 3, 4}); }).empty());  // no matrix yet: loud, not silently unrotated
            EXPECT(message_of([&] { a.setRotation({1, 0, 0, 1}); }) == "the rotation matrix must be d x d");
            a.setRotation(R);
            const std::vector<std::vector<double>> X = {{1, 2, 3, 4}, {4, 4, 0, 0}, {0, -4, 5, 1}, {3, 3, 3, -3}};
            for (size_t i = 0; i < X.size(); ++i) {
                EXPECT(a.indexVector("v" + std::to_string(i), X[i]));
                EXPECT(b.indexVector("v" + std::to_string(i), rot(X[i])));
            }
            const std::vector<double> q = {2, -3, 4, 1};
            Answer ra = a.computeNearestNeighbors(4, q), rb = b.computeNearestNeighbors(4, rot(q));
            EXPECT(ra.getIds() == rb.getIds() && ra.getDistances() == rb.getDistances() && ra.getIds().size() == 4);
        }
        const std::vector<double> cb = {0, 0, 10, 10};
        VladAggregator agg(cb, 2, 2);
        const std::vector<double> v = agg.aggregate({{1, 1}, {9, 9}, {2, 0}});
        EXPECT(v.size() == 4 && v[0] == 3 && v[1] == 1 && v[2] == -1 && v[3] == -1);
    }
    std::printf(fails ? "mirror_check: %d failure(s)\n" : "mirror_check ok\n", fails);
    return fails ? 1 : 0;
}
