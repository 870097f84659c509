// h5lite.hpp -- writes the HDF5 files of mio::save_result (io/result_io.h:44, io/result_io.cpp:33-100) without an
// HDF5 library. One file = one group per region id, each holding the datasets "Time" [T],
// "Group1".."GroupG" [T][C] and "Total" [T][C] of doubles.
//
// The container is the classic (version 0) HDF5 layout, the one libhdf5 writes by default and the one the reference's
// own result files use (cpp/tests/data/results_*.h5): superblock version 0, groups as symbol tables (one B-tree v1 node
// + local heap + symbol table node per group), object headers version 1, contiguous little-endian IEEE double
// datasets. The symbol table node size is a superblock parameter ("group leaf node K"); it is chosen so that every
// group fits one node, which keeps every B-tree at one level. tests/h5lite.py reads these files and the reference's.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace mio
{
namespace h5lite
{

struct Dataset {
    std::string name;
    std::vector<uint64_t> shape; // rank 1 or 2
    std::vector<double> values;  // row major
};
struct Group {
    std::string name;
    std::vector<Dataset> datasets;
};

class Writer
{
public:
    // the whole file in memory; every structure starts on an 8-byte boundary
    static std::vector<unsigned char> build(std::vector<Group> groups)
    {
        auto by_name = [](const auto& a, const auto& b) {
            return a.name < b.name; // symbol table entries are ordered by name (strcmp)
        };
        std::sort(groups.begin(), groups.end(), by_name);
        size_t max_entries = std::max<size_t>(groups.size(), 1);
        for (auto& g : groups) {
            std::sort(g.datasets.begin(), g.datasets.end(), by_name);
            max_entries = std::max(max_entries, g.datasets.size());
        }
        Writer w;
        w.m_leaf_k = uint16_t(std::max<size_t>(4, (max_entries + 1) / 2)); // a symbol table node holds 2K entries
        w.m_buf.assign(96, 0);                                             // superblock + root symbol table entry
        // datasets and region groups first (children before parents: a parent stores its children's addresses)
        std::vector<std::pair<std::string, Entry>> root_entries;
        for (auto& g : groups) {
            std::vector<std::pair<std::string, Entry>> entries;
            for (auto& d : g.datasets) {
                entries.emplace_back(d.name, Entry{w.dataset(d), 0, 0, 0});
            }
            root_entries.emplace_back(g.name, w.group(entries));
        }
        const Entry root = w.group(root_entries);
        w.superblock(root);
        return std::move(w.m_buf);
    }

private:
    struct Entry {
        uint64_t header; // object header address
        uint32_t cache;  // 1: group, scratch-pad = (B-tree, heap)
        uint64_t btree, heap;
    };
    static constexpr uint64_t UNDEF = ~uint64_t(0);
    std::vector<unsigned char> m_buf;
    uint16_t m_leaf_k = 4;

    uint64_t tell()
    {
        m_buf.resize((m_buf.size() + 7) & ~size_t(7), 0);
        return m_buf.size();
    }
    template <class T>
    void put(T v)
    {
        unsigned char b[sizeof(T)];
        std::memcpy(b, &v, sizeof(T)); // little-endian host
        m_buf.insert(m_buf.end(), b, b + sizeof(T));
    }
    void put_bytes(const void* p, size_t n)
    {
        const auto* b = static_cast<const unsigned char*>(p);
        m_buf.insert(m_buf.end(), b, b + n);
    }
    void pad8()
    {
        m_buf.resize((m_buf.size() + 7) & ~size_t(7), 0);
    }
    // header message: type, size of the (padded) body, flags
    void message(uint16_t type, uint8_t flags, const std::vector<unsigned char>& body)
    {
        const size_t padded = (body.size() + 7) & ~size_t(7);
        put<uint16_t>(type);
        put<uint16_t>(uint16_t(padded));
        put<uint8_t>(flags);
        put<uint8_t>(0);
        put<uint16_t>(0);
        put_bytes(body.data(), body.size());
        m_buf.resize(m_buf.size() + (padded - body.size()), 0);
    }
    // version 1 object header prefix; returns the position of the header-size field to patch
    uint64_t object_header(uint16_t num_messages)
    {
        const uint64_t at = tell();
        put<uint8_t>(1);
        put<uint8_t>(0);
        put<uint16_t>(num_messages);
        put<uint32_t>(1); // reference count
        put<uint32_t>(0); // header size, patched by end_header
        put<uint32_t>(0); // alignment
        return at;
    }
    void end_header(uint64_t at)
    {
        const uint32_t size = uint32_t(m_buf.size() - (at + 16));
        std::memcpy(&m_buf[at + 8], &size, 4);
    }

    uint64_t dataset(const Dataset& d)
    {
        // raw data, then the header that points to it
        const uint64_t data_at = tell();
        put_bytes(d.values.data(), d.values.size() * sizeof(double));
        const uint64_t at = object_header(4);
        {
            std::vector<unsigned char> b{1, uint8_t(d.shape.size()), 0, 0, 0, 0, 0, 0}; // dataspace version 1, simple
            for (uint64_t n : d.shape) {
                const auto* p = reinterpret_cast<const unsigned char*>(&n);
                b.insert(b.end(), p, p + 8);
            }
            message(0x0001, 0, b);
        }
        {
            // datatype: class 1 (floating point) version 1; little endian, IEEE layout: sign bit 63, exponent 52..62
            // (bias 1023), mantissa 0..51, implied leading bit
            const unsigned char b[] = {0x11, 0x20, 0x3f, 0x00, 8,    0,    0,    0,    0, 0, 64, 0,
                                       52,   11,   0,    52,   0xff, 0x03, 0x00, 0x00};
            message(0x0003, 1, std::vector<unsigned char>(b, b + sizeof b));
        }
        {
            // fill value version 2: allocate late, write if set, default fill value (size 0)
            const unsigned char b[] = {2, 2, 2, 1, 0, 0, 0, 0};
            message(0x0005, 1, std::vector<unsigned char>(b, b + sizeof b));
        }
        {
            std::vector<unsigned char> b{3, 1}; // layout version 3, contiguous
            const uint64_t size = d.values.size() * sizeof(double);
            const auto* pa      = reinterpret_cast<const unsigned char*>(&data_at);
            const auto* ps      = reinterpret_cast<const unsigned char*>(&size);
            b.insert(b.end(), pa, pa + 8);
            b.insert(b.end(), ps, ps + 8);
            message(0x0008, 0, b);
        }
        end_header(at);
        return at;
    }

    // a group of already written objects: local heap (names), symbol table node, B-tree node, object header
    Entry group(const std::vector<std::pair<std::string, Entry>>& entries)
    {
        // heap data segment: offset 0 holds the empty string, every name starts on an 8-byte boundary
        std::vector<unsigned char> names(8, 0);
        std::vector<uint64_t> name_off;
        for (auto& e : entries) {
            name_off.push_back(names.size());
            names.insert(names.end(), e.first.begin(), e.first.end());
            names.push_back(0);
            names.resize((names.size() + 7) & ~size_t(7), 0);
        }
        const uint64_t heap_at = tell();
        put_bytes("HEAP", 4);
        put<uint32_t>(0);              // version 0 + reserved
        put<uint64_t>(names.size());   // data segment size
        put<uint64_t>(1);              // head of the free list: none (H5HL_FREE_NULL)
        put<uint64_t>(heap_at + 32);   // data segment address
        put_bytes(names.data(), names.size());

        const uint64_t snod_at = tell();
        put_bytes("SNOD", 4);
        put<uint8_t>(1);
        put<uint8_t>(0);
        put<uint16_t>(uint16_t(entries.size()));
        for (size_t i = 0; i < size_t(2 * m_leaf_k); ++i) {
            if (i < entries.size()) {
                const Entry& e = entries[i].second;
                put<uint64_t>(name_off[i]);
                put<uint64_t>(e.header);
                put<uint32_t>(e.cache);
                put<uint32_t>(0);
                put<uint64_t>(e.cache == 1 ? e.btree : 0);
                put<uint64_t>(e.cache == 1 ? e.heap : 0);
            }
            else {
                m_buf.resize(m_buf.size() + 40, 0);
            }
        }

        const uint64_t tree_at = tell();
        const int K            = 16; // group internal node K (superblock)
        put_bytes("TREE", 4);
        put<uint8_t>(0); // node type: group
        put<uint8_t>(0); // level: its children are symbol table nodes
        put<uint16_t>(entries.empty() ? 0 : 1);
        put<uint64_t>(UNDEF); // siblings
        put<uint64_t>(UNDEF);
        // key 0 (the empty string), child 0, key 1 (the biggest name of child 0); the rest of the 2K+1 keys / 2K children
        put<uint64_t>(0);
        put<uint64_t>(entries.empty() ? UNDEF : snod_at);
        put<uint64_t>(entries.empty() ? 0 : name_off.back());
        m_buf.resize(m_buf.size() + size_t(2 * K - 1) * 16, 0);

        const uint64_t at = object_header(1);
        {
            std::vector<unsigned char> b;
            const auto* pt = reinterpret_cast<const unsigned char*>(&tree_at);
            const auto* ph = reinterpret_cast<const unsigned char*>(&heap_at);
            b.insert(b.end(), pt, pt + 8);
            b.insert(b.end(), ph, ph + 8);
            message(0x0011, 0, b); // symbol table message
        }
        end_header(at);
        return Entry{at, 1, tree_at, heap_at};
    }

    void superblock(const Entry& root)
    {
        const uint64_t eof = tell();
        std::vector<unsigned char> saved;
        saved.swap(m_buf);
        m_buf.clear();
        const unsigned char sig[] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        put_bytes(sig, 8);
        put<uint8_t>(0); // superblock version
        put<uint8_t>(0); // free-space storage version
        put<uint8_t>(0); // root group symbol table entry version
        put<uint8_t>(0);
        put<uint8_t>(0); // shared header message format version
        put<uint8_t>(8); // size of offsets
        put<uint8_t>(8); // size of lengths
        put<uint8_t>(0);
        put<uint16_t>(m_leaf_k); // group leaf node K
        put<uint16_t>(16);       // group internal node K
        put<uint32_t>(0);        // file consistency flags
        put<uint64_t>(0);        // base address
        put<uint64_t>(UNDEF);    // free-space info
        put<uint64_t>(eof);      // end of file address
        put<uint64_t>(UNDEF);    // driver information block
        put<uint64_t>(0);        // root entry: link name offset
        put<uint64_t>(root.header);
        put<uint32_t>(1);
        put<uint32_t>(0);
        put<uint64_t>(root.btree);
        put<uint64_t>(root.heap);
        std::copy(m_buf.begin(), m_buf.end(), saved.begin());
        m_buf.swap(saved);
    }
};

// Reader for the same subset of the format (what tests/h5lite.py does in Python): the files of save_result written by
// this writer or by libhdf5 (superblock 0, symbol table groups, version 1 object headers incl. continuation blocks,
// contiguous little-endian doubles). Throws std::runtime_error on anything else.
class Reader
{
public:
    explicit Reader(const std::string& filename)
    {
        std::ifstream in(filename, std::ios::binary);
        if (!in) {
            throw std::runtime_error("h5lite: cannot open " + filename);
        }
        m_b.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
        const unsigned char sig[] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        require(m_b.size() >= 96 && std::memcmp(m_b.data(), sig, 8) == 0, "not an HDF5 file");
        require(m_b[8] == 0 && m_b[13] == 8 && m_b[14] == 8, "superblock version 0 with 8-byte offsets expected");
        m_root = symbol_entry(56);
    }
    // names of the members of a group (root: empty path) in name order, with their symbol table entries
    struct Member {
        std::string name;
        uint64_t header;
        uint32_t cache;
        uint64_t btree, heap;
    };
    std::vector<Member> list(const Member* group = nullptr) const
    {
        const Member& g = group ? *group : m_root;
        uint64_t btree = g.btree, heap = g.heap;
        if (g.cache != 1) {
            bool found = false;
            for (auto& m : messages(g.header)) {
                if (m.type == 0x11) {
                    btree = get<uint64_t>(m.at);
                    heap  = get<uint64_t>(m.at + 8);
                    found = true;
                }
            }
            require(found, "not a group");
        }
        std::vector<Member> out;
        walk(btree, heap, out);
        return out;
    }
    // a dataset of doubles: shape and values (row major)
    Dataset dataset(const Member& m) const
    {
        Dataset d;
        d.name = m.name;
        bool type_ok = false, have_layout = false;
        uint64_t addr = 0, size = 0;
        for (auto& msg : messages(m.header)) {
            if (msg.type == 0x01) {
                const int ver = m_b.at(msg.at), rank = m_b.at(msg.at + 1);
                const uint64_t off = msg.at + (ver == 1 ? 8 : 4);
                for (int i = 0; i < rank; ++i) {
                    d.shape.push_back(get<uint64_t>(off + 8 * uint64_t(i)));
                }
            }
            else if (msg.type == 0x03) {
                type_ok = (m_b.at(msg.at) & 0x0f) == 1 && get<uint32_t>(msg.at + 4) == 8 && (m_b.at(msg.at + 1) & 1) == 0;
            }
            else if (msg.type == 0x08) {
                require(m_b.at(msg.at) == 3 && m_b.at(msg.at + 1) == 1, "contiguous layout (version 3) expected");
                addr        = get<uint64_t>(msg.at + 2);
                size        = get<uint64_t>(msg.at + 10);
                have_layout = true;
            }
        }
        require(type_ok && have_layout && !d.shape.empty(), "dataset of little-endian doubles expected");
        uint64_t n = 1;
        for (uint64_t k : d.shape) {
            n *= k;
        }
        require(size == 8 * n && addr + size <= m_b.size(), "dataset size mismatch");
        d.values.resize(size_t(n));
        std::memcpy(d.values.data(), m_b.data() + addr, size_t(size));
        return d;
    }

private:
    struct Message {
        uint16_t type;
        uint64_t at; // offset of the message body
        uint16_t size;
    };
    std::vector<unsigned char> m_b;
    Member m_root;

    static void require(bool ok, const char* what)
    {
        if (!ok) {
            throw std::runtime_error(std::string("h5lite: ") + what);
        }
    }
    template <class T>
    T get(uint64_t at) const
    {
        require(at + sizeof(T) <= m_b.size(), "truncated file");
        T v;
        std::memcpy(&v, m_b.data() + at, sizeof(T));
        return v;
    }
    Member symbol_entry(uint64_t at) const
    {
        Member m;
        m.header = get<uint64_t>(at + 8);
        m.cache  = get<uint32_t>(at + 16);
        m.btree  = m.cache == 1 ? get<uint64_t>(at + 24) : 0;
        m.heap   = m.cache == 1 ? get<uint64_t>(at + 32) : 0;
        m.name   = std::to_string(get<uint64_t>(at)); // placeholder: the heap offset, replaced by walk()
        return m;
    }
    std::vector<Message> messages(uint64_t header) const
    {
        require(m_b.at(header) == 1, "object header version 1 expected");
        const uint16_t count = get<uint16_t>(header + 2);
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{header + 16, get<uint32_t>(header + 8)}};
        std::vector<Message> out;
        for (size_t b = 0; b < blocks.size() && out.size() < count; ++b) {
            uint64_t at = blocks[b].first, left = blocks[b].second;
            while (left >= 8 && out.size() < count) {
                const Message m{get<uint16_t>(at), at + 8, get<uint16_t>(at + 2)};
                out.push_back(m);
                if (m.type == 0x10) {
                    blocks.emplace_back(get<uint64_t>(m.at), get<uint64_t>(m.at + 8));
                }
                at += 8 + m.size;
                left -= std::min<uint64_t>(left, 8 + uint64_t(m.size));
            }
        }
        return out;
    }
    void walk(uint64_t node, uint64_t heap, std::vector<Member>& out) const
    {
        require(std::memcmp(m_b.data() + node, "TREE", 4) == 0 && m_b.at(node + 4) == 0, "group B-tree node expected");
        const int level     = m_b.at(node + 5);
        const uint16_t used = get<uint16_t>(node + 6);
        require(std::memcmp(m_b.data() + heap, "HEAP", 4) == 0, "local heap expected");
        const uint64_t names = get<uint64_t>(heap + 24);
        for (uint16_t i = 0; i < used; ++i) {
            const uint64_t child = get<uint64_t>(node + 24 + 8 + 16 * uint64_t(i));
            if (level > 0) {
                walk(child, heap, out);
                continue;
            }
            require(std::memcmp(m_b.data() + child, "SNOD", 4) == 0, "symbol table node expected");
            const uint16_t n = get<uint16_t>(child + 6);
            for (uint16_t k = 0; k < n; ++k) {
                Member m = symbol_entry(child + 8 + 40 * uint64_t(k));
                const uint64_t s = names + get<uint64_t>(child + 8 + 40 * uint64_t(k));
                m.name = std::string(reinterpret_cast<const char*>(m_b.data() + s));
                out.push_back(std::move(m));
            }
        }
    }
};

inline bool write_file(const std::string& filename, const std::vector<Group>& groups)
{
    const auto bytes = Writer::build(groups);
    std::ofstream out(filename, std::ios::binary);
    if (!out) {
        return false;
    }
    out.write(reinterpret_cast<const char*>(bytes.data()), std::streamsize(bytes.size()));
    return bool(out);
}

} // namespace h5lite
} // namespace mio
