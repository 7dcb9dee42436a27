// host/repartitor.hpp -- the k-mer -> partition function of a reference run, read from its <out-tmp>/simka_output_temp/config.h5.
//
// The reference's orchestrator writes config.h5 before it starts the counting processes (S/SimkaPotara.hpp:573-606,722-767): the DSK
// Configuration and the Repartitor table that assigns every minimizer to one of the partitions (GC/kmer/impl/PartiInfo.cpp:228-295,
// RepartitionAlgorithm.cpp).  Its simkaCount sends a solid k-mer to solid/part_<table[minimizer(kmer)]>.  With the same function this
// build's simkaCount writes partition files that hold exactly the records the reference's would, so the two implementations can be
// mixed per SAMPLE in one temp directory (e.g. new samples counted here, merged with samples the reference counted earlier).
//
// config.h5 as GATB's StorageHDF5 writes it is a small, regular HDF5 file: superblock version 0, the datasets ("minimRepart",
// "config") directly in the root group (one symbol-table node behind a version-1 B-tree, names in a local heap), version-1 object
// headers, one-dimensional byte datasets in 4096-byte chunks addressed by a version-1 chunk B-tree, no filters.  H5Min reads exactly
// that family (and refuses anything else); it is not a general HDF5 reader and writes nothing.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace simka_b200 {

class H5Min {
public:
    explicit H5Min(const std::string &path) {
        std::ifstream in(path.c_str(), std::ios::binary);
        if (!in) throw std::runtime_error("cannot open " + path);
        raw_.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
        static const unsigned char sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        if (raw_.size() < 96 || memcmp(raw_.data(), sig, 8) != 0) fail("not an HDF5 file");
        if (u8(8) != 0) fail("superblock version");
        if (u8(13) != 8 || u8(14) != 8) fail("offset / length size");
        // root group symbol-table entry at 56: link name offset, object header, cache type 1 with B-tree and heap addresses
        if (u32(56 + 16) != 1) fail("root group without cached symbol table");
        const uint64_t btree = u64(56 + 24), heap = u64(56 + 32);
        if (memcmp(at(heap, 4), "HEAP", 4) != 0) fail("local heap");
        heap_data_ = u64(heap + 24);
        walk_group(btree, 0);
    }
    bool has(const std::string &name) const { return objects_.count(name) != 0; }
    // the bytes of a one-dimensional, chunked, unfiltered dataset of 1-byte elements
    std::vector<uint8_t> dataset(const std::string &name) const {
        auto it = objects_.find(name);
        if (it == objects_.end()) fail("no dataset '" + name + "'");
        const uint64_t oh = it->second;
        if (u8(oh) != 1) fail("object header version");
        uint64_t size = 0, chunk_tree = 0;
        uint32_t ndim = 0, elem = 0;
        bool have_space = false, have_layout = false;
        std::vector<std::pair<uint64_t, uint64_t>> blocks(1, std::make_pair(oh + 16, oh + 16 + (uint64_t)u32(oh + 8)));
        for (size_t b = 0; b < blocks.size() && b < 64; b++) {
            uint64_t p = blocks[b].first;
            const uint64_t end = blocks[b].second;
            while (p + 8 <= end) {
                const uint32_t type = u16(p), sz = u16(p + 2);
                const uint64_t body = p + 8;
                at(body, sz);
                if (type == 0x10) blocks.push_back(std::make_pair(u64(body), u64(body) + u64(body + 8)));  // continuation
                else if (type == 0x0B) fail("filtered dataset");
                else if (type == 0x01 && !have_space) {  // dataspace: version 1, rank 1
                    if (u8(body) != 1 || u8(body + 1) != 1) fail("dataspace");
                    size = u64(body + 8);
                    have_space = true;
                } else if (type == 0x03 && elem == 0) elem = u32(body + 4);  // datatype: element size
                else if (type == 0x08 && !have_layout) {  // layout: version 3, chunked
                    if (u8(body) != 3 || u8(body + 1) != 2) fail("layout");
                    ndim = u8(body + 2);
                    chunk_tree = u64(body + 3);
                    have_layout = true;
                }
                p += 8 + sz;
            }
        }
        if (!have_space || !have_layout || elem != 1 || ndim != 2 || size > ((uint64_t)1 << 28)) fail("dataset '" + name + "' is not a chunked byte array");
        std::vector<uint8_t> out((size_t)size, 0);
        walk_chunks(chunk_tree, ndim, out, 0);
        return out;
    }

private:
    [[noreturn]] static void fail(const std::string &what) { throw std::runtime_error("config.h5: unsupported or damaged file (" + what + ")"); }
    const uint8_t *at(uint64_t off, uint64_t n) const {
        if (off > raw_.size() || n > raw_.size() - off) fail("offset out of range");
        return reinterpret_cast<const uint8_t *>(raw_.data()) + off;
    }
    uint8_t u8(uint64_t o) const { return *at(o, 1); }
    uint16_t u16(uint64_t o) const { uint16_t v; memcpy(&v, at(o, 2), 2); return v; }
    uint32_t u32(uint64_t o) const { uint32_t v; memcpy(&v, at(o, 4), 4); return v; }
    uint64_t u64(uint64_t o) const { uint64_t v; memcpy(&v, at(o, 8), 8); return v; }
    // version-1 B-tree, node type 0 (group): keys are heap offsets (8 bytes), children are sub-nodes or symbol-table nodes
    void walk_group(uint64_t node, int depth) {
        if (depth > 8 || memcmp(at(node, 4), "TREE", 4) != 0 || u8(node + 4) != 0) fail("group B-tree");
        const int level = u8(node + 5), used = u16(node + 6);
        uint64_t p = node + 8 + 16;
        for (int i = 0; i < used; i++) {
            p += 8;
            const uint64_t child = u64(p);
            p += 8;
            if (level > 0) { walk_group(child, depth + 1); continue; }
            if (memcmp(at(child, 4), "SNOD", 4) != 0) fail("symbol-table node");
            const int nsym = u16(child + 6);
            for (int j = 0; j < nsym; j++) {
                const uint64_t e = child + 8 + 40 * (uint64_t)j;
                const char *s = reinterpret_cast<const char *>(at(heap_data_ + u64(e), 1));
                const size_t room = raw_.size() - (size_t)(heap_data_ + u64(e));
                objects_[std::string(s, strnlen(s, room))] = u64(e + 8);
            }
        }
    }
    // version-1 B-tree, node type 1 (chunks): key = chunk size, filter mask, ndim offsets
    void walk_chunks(uint64_t node, uint32_t ndim, std::vector<uint8_t> &out, int depth) const {
        if (depth > 8 || memcmp(at(node, 4), "TREE", 4) != 0 || u8(node + 4) != 1) fail("chunk B-tree");
        const int level = u8(node + 5), used = u16(node + 6);
        uint64_t p = node + 8 + 16;
        for (int i = 0; i < used; i++) {
            const uint32_t csize = u32(p), fmask = u32(p + 4);
            const uint64_t off = u64(p + 8);
            p += 8 + 8 * (uint64_t)ndim;
            const uint64_t child = u64(p);
            p += 8;
            if (level > 0) { walk_chunks(child, ndim, out, depth + 1); continue; }
            if (fmask != 0) fail("filtered chunk");
            if (off >= out.size()) continue;
            const uint64_t n = std::min<uint64_t>(csize, out.size() - off);
            memcpy(out.data() + off, at(child, n), (size_t)n);
        }
    }
    std::string raw_;
    uint64_t heap_data_ = 0;
    std::map<std::string, uint64_t> objects_;
};

// Repartitor::load (GC/kmer/impl/PartiInfo.cpp:228-261): u16 nbpart, u64 nb_minims, u16 nbPass, u16 table[nb_minims], bool, u32 magic
struct Repartitor {
    uint16_t nbpart = 0, nbpass = 0;
    uint64_t nb_minims = 0;
    int minimizer_size = 0;
    std::vector<uint16_t> table;
    static Repartitor load(const std::string &config_h5) {
        const std::vector<uint8_t> b = H5Min(config_h5).dataset("minimRepart");
        Repartitor r;
        if (b.size() < 12) throw std::runtime_error("config.h5: minimRepart too short");
        memcpy(&r.nbpart, b.data(), 2);
        memcpy(&r.nb_minims, b.data() + 2, 8);
        memcpy(&r.nbpass, b.data() + 10, 2);
        if (r.nb_minims == 0 || r.nb_minims > ((uint64_t)1 << 30) || b.size() < 12 + 2 * r.nb_minims + 5) throw std::runtime_error("config.h5: minimRepart has a bad size");
        uint32_t magic = 0;
        memcpy(&magic, b.data() + 12 + 2 * r.nb_minims + 1, 4);
        if (magic != 0x12345678u) throw std::runtime_error("config.h5: minimRepart has a bad magic number");
        r.table.resize((size_t)r.nb_minims);
        memcpy(r.table.data(), b.data() + 12, 2 * (size_t)r.nb_minims);
        while (((uint64_t)1 << (2 * r.minimizer_size)) < r.nb_minims) r.minimizer_size++;
        if (((uint64_t)1 << (2 * r.minimizer_size)) != r.nb_minims) throw std::runtime_error("config.h5: the number of minimizers is not a power of four");
        for (uint16_t p : r.table) if (p >= r.nbpart) throw std::runtime_error("config.h5: partition out of range in the repartition table");
        return r;
    }
};

}  // namespace simka_b200
