// Exercises include/hannoy_b200.hpp (the C++ host mirror of hannoy's Reader / QueryBuilder).
//   host_mirror_test cpu           error paths that need no device (Reader::open checks, reader.rs:390-416)
//   host_mirror_test gpu <file>    open from raw KV pairs, search, compare with the expected results in <file>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "hannoy_b200.hpp"

using namespace hannoy;
using KV = std::vector<std::pair<std::string, std::string>>;

template <class T>
static T rd(std::ifstream& f) { T v; f.read((char*)&v, sizeof(T)); return v; }

template <class D>
static int run_gpu(std::ifstream& f, uint16_t index) {
    uint32_t n_kv = rd<uint32_t>(f);
    KV kv(n_kv);
    for (auto& [k, v] : kv) {
        k.resize(rd<uint32_t>(f)); f.read(k.data(), k.size());
        v.resize(rd<uint32_t>(f)); f.read(v.data(), v.size());
    }
    uint32_t dims = rd<uint32_t>(f), nq = rd<uint32_t>(f), k = rd<uint32_t>(f), ef = rd<uint32_t>(f);
    std::vector<float> q((size_t)nq * dims);
    f.read((char*)q.data(), q.size() * 4);
    std::vector<uint32_t> lens(nq), ids((size_t)nq * k);
    std::vector<float> dist((size_t)nq * k);
    f.read((char*)lens.data(), lens.size() * 4);
    f.read((char*)ids.data(), ids.size() * 4);
    f.read((char*)dist.data(), dist.size() * 4);

    Reader<D> reader = Reader<D>::open(kv, index, 0);
    if (reader.dimensions() != dims) { std::printf("FAIL dimensions\n"); return 1; }
    auto res = reader.nns(k).ef_search(ef).by_vectors(q.data(), nq, dims);
    for (uint32_t i = 0; i < nq; ++i) {
        if (res[i].nns.size() != lens[i]) { std::printf("FAIL len of query %u\n", i); return 1; }
        for (uint32_t j = 0; j < lens[i]; ++j) {
            if (res[i].nns[j].first != ids[(size_t)i * k + j] || std::memcmp(&res[i].nns[j].second, &dist[(size_t)i * k + j], 4)) {
                std::printf("FAIL query %u rank %u\n", i, j);
                return 1;
            }
        }
    }
    // by_vector with a wrong dimension -> InvalidVecDimension (reader.rs:133-138)
    try {
        reader.nns(k).by_vector(std::vector<float>(dims + 1, 0.f));
        std::printf("FAIL no InvalidVecDimension\n");
        return 1;
    } catch (const InvalidVecDimension&) {}
    // by_item: absent item -> None (reader.rs:826); present item never returns itself
    ItemId present = reader.item_ids()[0];
    auto r1 = reader.nns(3).by_item(0xfffffff0u);
    auto r2 = reader.nns(3).by_item(present);
    if (r1.has_value() || !r2.has_value()) { std::printf("FAIL by_item option\n"); return 1; }
    for (auto& p : r2->nns) if (p.first == present) { std::printf("FAIL by_item contains item\n"); return 1; }
    std::printf("OK %u queries\n", nq);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "cpu")) {
        try {
            Reader<distances::Cosine>::open(KV{}, 0);
            std::printf("FAIL empty KV opened\n");
            return 1;
        } catch (const MissingMetadata&) {}
        if (std::strcmp(distances::name<distances::BinaryQuantizedCosine>(), "binary quantized cosine")) { std::printf("FAIL name\n"); return 1; }
        std::printf("OK cpu\n");
        return 0;
    }
    if (argc >= 3 && !std::strcmp(argv[1], "gpu")) {
        std::ifstream f(argv[2], std::ios::binary);
        uint32_t metric = rd<uint32_t>(f), index = rd<uint32_t>(f);
        try {
            switch (metric) {
                case HB_EUCLIDEAN: return run_gpu<distances::Euclidean>(f, (uint16_t)index);
                case HB_COSINE: return run_gpu<distances::Cosine>(f, (uint16_t)index);
                case HB_HAMMING: return run_gpu<distances::Hamming>(f, (uint16_t)index);
                case HB_BQ_COSINE: return run_gpu<distances::BinaryQuantizedCosine>(f, (uint16_t)index);
                default: std::printf("FAIL metric\n"); return 1;
            }
        } catch (const Error& e) {
            std::printf("FAIL exception %d: %s\n", (int)e.status, e.what());
            return 1;
        }
    }
    std::printf("usage: host_mirror_test cpu | gpu <file>\n");
    return 2;
}
