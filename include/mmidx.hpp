// mmidx.hpp -- C++ host-side mirror of the reference's Java classes over the C ABI of libmmidx.so (mmidx.h).
//
// The reference's host language is Java and this image has no JDK, so besides the Python mirror used by the tests
// (multimedia-indexing_b200/datastructures.py) the same surface is given here for a compiled host: same class and
// method names, argument meaning and error behaviour as gr.iti.mklab.visual.datastructures.{AbstractSearchStructure,
// Linear, PQ, IVFPQ}, gr.iti.mklab.visual.utilities.Answer and gr.iti.mklab.visual.aggregation.VladAggregator
// (J/ = src/main/java/gr/iti/mklab/visual/).  Checked exceptions become mmidx::Exception(message); the id <-> internal
// id maps the reference keeps in BDB JE stay on the host side as in-memory maps.  Header-only; link with -lmmidx.
//
// Checked by tests/test_cpp_mirror.py, which builds tests/cpp/mirror_check.cpp: error paths and the no-device failure on
// CPU, a Linear / VladAggregator round trip on a B200.
#pragma once
#include <chrono>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "mmidx.h"

namespace mmidx {

// the reference throws java.lang.Exception(message); `code` is the C status it came from
class Exception : public std::runtime_error {
public:
    int code;
    Exception(int c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

inline void check(int rc) {
    if (rc != MMIDX_OK) throw Exception(rc, mmidx_last_error());
}

enum class TransformationType { None, RandomRotation, RandomPermutation };  // J/datastructures/PQ.java:60-62

// RandomPermutation.java:29-40: Collections.shuffle(0..dim-1, new java.util.Random(seed)), restated exactly
inline std::vector<int32_t> random_permutation(int64_t seed, int dim) {
    const uint64_t mask = (1ULL << 48) - 1;
    uint64_t s = ((uint64_t)seed ^ 0x5DEECE66DULL) & mask;  // java.util.Random(seed)
    auto next31 = [&]() {                                   // Random.next(31)
        s = (s * 0x5DEECE66DULL + 0xBULL) & mask;
        return (int32_t)(s >> 17);
    };
    auto next_int = [&](int32_t bound) {  // Random.nextInt(bound)
        int32_t r = next31();
        const int32_t m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        for (int32_t u = r;; u = next31()) {
            r = u % bound;
            if ((int64_t)u - r + m <= 0x7fffffffLL) break;  // Java: u - r + m < 0 on int overflow -> draw again
        }
        return r;
    };
    std::vector<int32_t> p((size_t)dim);
    for (int i = 0; i < dim; ++i) p[(size_t)i] = i;
    for (int i = dim; i > 1; --i) std::swap(p[(size_t)i - 1], p[(size_t)next_int(i)]);  // Collections.shuffle
    return p;
}

// J/utilities/Answer.java
class Answer {
    std::vector<std::string> ids_;
    std::vector<double> distances_;
    long nameLookupTime_, indexSearchTime_;

public:
    Answer(std::vector<std::string> ids, std::vector<double> distances, long nameLookupTime, long indexSearchTime)
        : ids_(std::move(ids)), distances_(std::move(distances)), nameLookupTime_(nameLookupTime), indexSearchTime_(indexSearchTime) {}
    const std::vector<std::string> &getIds() const { return ids_; }
    const std::vector<double> &getDistances() const { return distances_; }  // squared L2 (ADC for PQ / IVFPQ), ascending
    long getNameLookupTime() const { return nameLookupTime_; }
    long getIndexSearchTime() const { return indexSearchTime_; }
};

// reads the reference's CSV codebooks: one centroid per line, comma separated.  skip_headers: lines without a comma are
// skipped (AbstractFeatureAggregator.readQuantizer AFA.java:234-254: coarse quantizer, VLAD codebook); the product
// quantizer is read line by line with no skipping (PQ.loadProductQuantizer PQ.java:210-223, IVFPQ.java:275-292), which
// matters when subVectorLength == 1 and no line holds a comma.
inline std::vector<double> read_csv_rows(const std::string &file, size_t rows, size_t cols, bool skip_headers = true) {
    std::ifstream in(file);
    if (!in) throw Exception(MMIDX_ERR_INVALID, "cannot open " + file);
    std::vector<double> out;
    out.reserve(rows * cols);
    std::string line;
    size_t got = 0;
    while (got < rows && std::getline(in, line)) {
        if (skip_headers && line.find(',') == std::string::npos) continue;
        std::stringstream ss(line);
        std::string tok;
        size_t c = 0;
        while (c < cols && std::getline(ss, tok, ',')) {
            out.push_back(std::stod(tok));
            ++c;
        }
        if (c != cols) throw Exception(MMIDX_ERR_DIM, "codebook line " + std::to_string(got) + " has the wrong length");
        ++got;
    }
    if (got != rows) throw Exception(MMIDX_ERR_INVALID, "codebook file holds fewer centroids than expected");
    return out;
}

// J/datastructures/AbstractSearchStructure.java
class AbstractSearchStructure {
protected:
    mmidx_t *h_ = nullptr;
    int vectorLength;
    int64_t maxNumVectors;
    std::unordered_map<std::string, int> nameToId_;  // the reference's "nameToId" / "idToName" BDB databases
    std::vector<std::string> idToName_;
    bool rotation_pending_ = false;  // PQ / IVFPQ built with RandomRotation wait for setRotation()
    void require_ready() const {
        if (rotation_pending_) throw Exception(MMIDX_ERR_UNSUPPORTED, "RandomRotation: supply the rotation matrix with setRotation() first");
    }

    AbstractSearchStructure(int vectorLength_, int64_t maxNumVectors_) : vectorLength(vectorLength_), maxNumVectors(maxNumVectors_) {}
    void create(const mmidx_params &p) { check(mmidx_create(&p, &h_)); }

public:
    AbstractSearchStructure(const AbstractSearchStructure &) = delete;
    AbstractSearchStructure &operator=(const AbstractSearchStructure &) = delete;
    virtual ~AbstractSearchStructure() { close(); }

    // ASS.java:734-755
    void close() {
        if (h_) mmidx_destroy(h_);
        h_ = nullptr;
    }
    int getLoadCounter() const {  // ASS.java:711
        int64_t n = 0;
        check(mmidx_size(h_, &n));
        return (int)n;
    }
    bool isIndexed(const std::string &id) const { return nameToId_.count(id) != 0; }
    int getInternalId(const std::string &id) const {  // ASS.java:406-421: -1 when absent
        auto it = nameToId_.find(id);
        return it == nameToId_.end() ? -1 : it->second;
    }
    const std::string &getId(int iid) const {  // ASS.java:380-395
        if (iid < 0 || (size_t)iid >= idToName_.size()) throw Exception(MMIDX_ERR_INVALID, "Internal id " + std::to_string(iid) + " is out of range!");
        return idToName_[iid];
    }

    // ASS.java:229-257: false (not an exception) when the index is full or the id is already indexed
    bool indexVector(const std::string &id, const std::vector<double> &vector) {
        if ((int64_t)idToName_.size() >= maxNumVectors) return false;
        if (isIndexed(id)) return false;
        if ((int)vector.size() != vectorLength) throw Exception(MMIDX_ERR_DIM, "The dimensionality of the vector is wrong!");
        require_ready();
        const int rc = mmidx_add(h_, 1, vector.data(), nullptr, nullptr);
        if (rc == MMIDX_ERR_FULL) return false;
        check(rc);
        nameToId_.emplace(id, (int)idToName_.size());
        idToName_.push_back(id);
        return true;
    }

    // ASS.java:281-291
    Answer computeNearestNeighbors(int k, const std::vector<double> &queryVector) {
        if ((int)queryVector.size() != vectorLength) throw Exception(MMIDX_ERR_DIM, "The dimensionality of the vector is wrong!");
        require_ready();
        std::vector<int32_t> iids((size_t)(k > 0 ? k : 0));
        std::vector<double> dist(iids.size());
        int32_t cnt = 0;
        const auto t0 = std::chrono::steady_clock::now();
        check(mmidx_search(h_, 1, queryVector.data(), k, iids.data(), dist.data(), &cnt));
        const auto t1 = std::chrono::steady_clock::now();
        std::vector<std::string> ids;
        for (int i = 0; i < cnt; ++i) ids.push_back(getId(iids[i]));  // lookUp ASS.java:345-357
        dist.resize((size_t)cnt);
        const auto t2 = std::chrono::steady_clock::now();
        return Answer(std::move(ids), std::move(dist), (long)std::chrono::duration_cast<std::chrono::milliseconds>(t2 - t1).count(),
                      (long)std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count());
    }

    // ASS.java:320-328: query by the id of an indexed vector
    virtual Answer computeNearestNeighbors(int /*k*/, const std::string &id) {
        if (!isIndexed(id)) throw Exception(MMIDX_ERR_INVALID, "Id does not exist!");
        throw Exception(MMIDX_ERR_UNSUPPORTED, "query-by-id is not available for this index type");
    }
};

// J/datastructures/Linear.java
class Linear : public AbstractSearchStructure {
public:
    Linear(int vectorLength_, int64_t maxNumVectors_, int device = -1) : AbstractSearchStructure(vectorLength_, maxNumVectors_) {
        mmidx_params p{};
        p.type = MMIDX_LINEAR;
        p.d = vectorLength_;
        p.max_n = maxNumVectors_;
        p.device = device;
        create(p);
    }
    std::vector<double> getVector(int iid) {  // Linear.java:253-281
        std::vector<double> v((size_t)vectorLength);
        check(mmidx_get_vector(h_, iid, v.data()));
        return v;
    }
    using AbstractSearchStructure::computeNearestNeighbors;
    Answer computeNearestNeighbors(int k, const std::string &id) override {  // Linear.java:181-184
        const int iid = getInternalId(id);
        if (iid == -1) throw Exception(MMIDX_ERR_INVALID, "Id does not exist!");
        return computeNearestNeighbors(k, getVector(iid));
    }
};

// J/datastructures/PQ.java
class PQ : public AbstractSearchStructure {
protected:
    int numSubVectors, numProductCentroids, subVectorLength;

    PQ(int vectorLength_, int64_t maxNumVectors_, int m, int ks, bool) : AbstractSearchStructure(vectorLength_, maxNumVectors_),
                                                                        numSubVectors(m), numProductCentroids(ks) {
        if (m <= 0 || vectorLength_ % m != 0) throw Exception(MMIDX_ERR_DIM, "The given number of subvectors is not valid!");  // PQ.java:148-150
        subVectorLength = vectorLength_ / m;
    }
    void set_transformation(TransformationType t, int seed) {
        // RandomRotation: EJML's generator is not reproducible offline; the constructor leaves the index
        // untransformed until setRotation() supplies the matrix -- indexing before that would silently differ, so it throws
        if (t == TransformationType::RandomRotation) rotation_pending_ = true;
        if (t == TransformationType::RandomPermutation) {
            const std::vector<int32_t> perm = random_permutation(seed, vectorLength);
            check(mmidx_set_permutation(h_, perm.data()));
        }
    }

public:
    // TransformationType.RandomRotation with the d x d matrix of RandomRotation.java:30-35 supplied (row-major):
    // transformed = v R, as RandomRotation.rotate (RandomRotation.java:44-49)
    void setRotation(const std::vector<double> &R) {
        if ((int64_t)R.size() != (int64_t)vectorLength * vectorLength) throw Exception(MMIDX_ERR_DIM, "the rotation matrix must be d x d");
        check(mmidx_set_transform(h_, MMIDX_TRANSFORM_ROTATION, nullptr, R.data()));
        rotation_pending_ = false;
    }
    bool rotationPending() const { return rotation_pending_; }

    // PQ.java:142-144 (without the BDB arguments); seed = 1 as in PQ.java:160
    PQ(int vectorLength_, int64_t maxNumVectors_, int numSubVectors_, int numProductCentroids_,
       TransformationType transformation = TransformationType::None, int device = -1)
        : PQ(vectorLength_, maxNumVectors_, numSubVectors_, numProductCentroids_, true) {
        mmidx_params p{};
        p.type = MMIDX_PQ;
        p.d = vectorLength_;
        p.max_n = maxNumVectors_;
        p.m = numSubVectors_;
        p.ks = numProductCentroids_;
        p.device = device;
        create(p);
        set_transformation(transformation, 1);
    }
    // PQ.java:210-223: m*ks lines of subVectorLength values
    void loadProductQuantizer(const std::string &filename) {
        loadProductQuantizer(read_csv_rows(filename, (size_t)numSubVectors * numProductCentroids, (size_t)subVectorLength, false));
    }
    void loadProductQuantizer(const std::vector<double> &P /* [m][ks][subVectorLength] */) {
        if (P.size() != (size_t)numSubVectors * numProductCentroids * subVectorLength)
            throw Exception(MMIDX_ERR_DIM, "product quantizer has the wrong shape");
        check(mmidx_set_product_quantizer(h_, P.data()));
    }
};

// J/datastructures/IVFPQ.java
class IVFPQ : public PQ {
    int numCoarseCentroids;

public:
    // IVFPQ.java:174-177 (without the BDB arguments); w defaults to (int)(numCoarseCentroids * 0.1), IVFPQ.java:188
    IVFPQ(int vectorLength_, int64_t maxNumVectors_, int numSubVectors_, int numProductCentroids_, TransformationType transformation,
          int numCoarseCentroids_, int device = -1)
        : PQ(vectorLength_, maxNumVectors_, numSubVectors_, numProductCentroids_, true), numCoarseCentroids(numCoarseCentroids_) {
        mmidx_params p{};
        p.type = MMIDX_IVFPQ;
        p.d = vectorLength_;
        p.max_n = maxNumVectors_;
        p.m = numSubVectors_;
        p.ks = numProductCentroids_;
        p.nlist = numCoarseCentroids_;
        p.w = 0;
        p.device = device;
        create(p);
        set_transformation(transformation, 1);
    }
    void setW(int w) { check(mmidx_set_w(h_, w)); }  // IVFPQ.java:95-97
    void loadCoarseQuantizer(const std::string &filename) {  // IVFPQ.java:297-300
        loadCoarseQuantizer(read_csv_rows(filename, (size_t)numCoarseCentroids, (size_t)vectorLength));
    }
    void loadCoarseQuantizer(const std::vector<double> &C /* [numCoarseCentroids][vectorLength] */) {
        if (C.size() != (size_t)numCoarseCentroids * vectorLength) throw Exception(MMIDX_ERR_DIM, "coarse quantizer has the wrong shape");
        check(mmidx_set_coarse_quantizer(h_, C.data()));
    }
    // IVFPQ.java:357-386: the Java signature takes signed bytes (code - 128)
    bool indexPQCode(const std::string &id, int listId, const std::vector<int8_t> &code) {
        if (numProductCentroids > 256) throw Exception(MMIDX_ERR_INVALID, "Call the short variant of the method!");  // IVFPQ.java:358-361
        if ((int64_t)idToName_.size() >= maxNumVectors || isIndexed(id)) return false;
        if ((int)code.size() != numSubVectors) throw Exception(MMIDX_ERR_DIM, "The dimensionality of the code is wrong!");
        std::vector<uint8_t> raw(code.size());
        for (size_t i = 0; i < code.size(); ++i) raw[i] = (uint8_t)((int)code[i] + 128);
        const int32_t l = listId;
        const int rc = mmidx_add_codes(h_, 1, &l, raw.data());
        if (rc == MMIDX_ERR_FULL) return false;
        check(rc);
        nameToId_.emplace(id, (int)idToName_.size());
        idToName_.push_back(id);
        return true;
    }
    std::vector<int32_t> computeNearestCoarseIndices(const std::vector<double> &vector, int k) {  // IVFPQ.java:575-601
        if ((int)vector.size() != vectorLength) throw Exception(MMIDX_ERR_DIM, "The dimensionality of the vector is wrong!");
        std::vector<int32_t> out((size_t)(k > 0 ? k : 0));
        check(mmidx_coarse_probe(h_, 1, vector.data(), k, out.data()));
        return out;
    }
    std::vector<int32_t> outputItemsPerList() {  // IVFPQ.java:654-673
        std::vector<int32_t> out((size_t)numCoarseCentroids);
        check(mmidx_list_sizes(h_, out.data()));
        return out;
    }
    // IVFPQ.java:393-396, 509-511: the reference's IVFSDC query-by-id returns null
    Answer computeNearestNeighbors(int, const std::string &) override {
        throw Exception(MMIDX_ERR_UNSUPPORTED, "IVFPQ query-by-id (computeKnnIVFSDC) returns null in the reference");
    }
    using AbstractSearchStructure::computeNearestNeighbors;
};

// J/aggregation/VladAggregator.java over AbstractFeatureAggregator
class VladAggregator {
    std::vector<double> codebook_;
    int numCentroids, descriptorLength, device_;

public:
    VladAggregator(std::vector<double> codebook /* [numCentroids][descriptorLength] */, int numCentroids_, int descriptorLength_, int device = -1)
        : codebook_(std::move(codebook)), numCentroids(numCentroids_), descriptorLength(descriptorLength_), device_(device) {
        if (codebook_.size() != (size_t)numCentroids * descriptorLength) throw Exception(MMIDX_ERR_DIM, "codebook has the wrong shape");
    }
    int getVectorLength() const { return numCentroids * descriptorLength; }  // VladAggregator.getVectorLength
    int getNumCentroids() const { return numCentroids; }
    int getDescriptorLength() const { return descriptorLength; }
    // AbstractFeatureAggregator.aggregate(double[][]) AFA.java:72-79; an empty set gives the zero vector (VladAggregator.java:57-59)
    std::vector<double> aggregate(const std::vector<std::vector<double>> &descriptors) const {
        std::vector<double> flat;
        flat.reserve(descriptors.size() * (size_t)descriptorLength);
        for (const auto &dsc : descriptors) {
            if ((int)dsc.size() != descriptorLength)
                throw Exception(MMIDX_ERR_DIM, "Descriptor length does not match codebook centroid length");  // AFA.java:74-76
            flat.insert(flat.end(), dsc.begin(), dsc.end());
        }
        const int64_t offsets[2] = {0, (int64_t)descriptors.size()};
        std::vector<double> out((size_t)getVectorLength());
        check(mmidx_vlad(codebook_.data(), numCentroids, descriptorLength, 1, offsets, flat.data(), out.data(), nullptr, device_));
        return out;
    }
};

}  // namespace mmidx
