// H5Lite implementation -- see H5Lite.h.  Structures follow the HDF5 File Format Specification:
// superblock (versions 0-3), object headers (versions 1 and 2), symbol-table groups (B-tree v1 node type 0,
// SNOD, local heap) and compact new-style groups (link messages), dataspace / datatype / layout messages.
#include "../include/H5Lite.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace H5Lite {

namespace {

const uint8_t kSignature[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
const uint64_t kUndef = 0xFFFFFFFFFFFFFFFFull;

struct Cursor {
    const std::vector<uint8_t> &b;
    uint64_t p;
    Cursor(const std::vector<uint8_t> &buf, uint64_t at) : b(buf), p(at) {}
    void need(uint64_t n) const {
        if (p + n > b.size()) throw Error("unexpected end of file at offset " + std::to_string(p));
    }
    uint64_t u(int n) { // little-endian unsigned of n bytes
        need(n);
        uint64_t v = 0;
        for (int k = 0; k < n; ++k) v |= (uint64_t)b[p + k] << (8 * k);
        p += n;
        return v;
    }
    void skip(uint64_t n) {
        need(n);
        p += n;
    }
};

std::string normName(const std::string &n) { return (!n.empty() && n[0] == '/') ? n.substr(1) : n; }

struct Message {
    int type;
    uint64_t at, size; // data location
};

} // namespace

// ------------------------------------------------------------------------------------------------
// reader
// ------------------------------------------------------------------------------------------------
File::File(const std::string &path) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw Error("cannot open '" + path + "'");
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz : 0);
    if (sz > 0 && std::fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) {
        std::fclose(f);
        throw Error("short read on '" + path + "'");
    }
    std::fclose(f);
    // superblock: at 0, 512, 1024, ... (user block)
    uint64_t sb = kUndef;
    for (uint64_t at = 0; at + 8 <= buf.size(); at = at ? at * 2 : 512)
        if (std::memcmp(buf.data() + at, kSignature, 8) == 0) {
            sb = at;
            break;
        }
    if (sb == kUndef) throw Error("'" + path + "' is not an HDF5 file (no signature)");
    Cursor c(buf, sb + 8);
    int version = (int)c.u(1);
    if (version == 0 || version == 1) {
        c.skip(4); // free-space version, root group version, reserved, shared header version
        int so = (int)c.u(1), sl = (int)c.u(1);
        if (so != 8 || sl != 8) throw Error("only 8-byte offsets/lengths are supported");
        c.skip(1);
        c.skip(4); // group leaf / internal node K
        c.skip(4); // consistency flags
        if (version == 1) c.skip(4);
        base = c.u(8);
        if (base == kUndef) base = 0;
        if (base == 0 && sb != 0) base = sb; // some writers store 0 and mean "relative to the superblock"
        c.skip(8);                           // free-space info
        c.skip(8);                           // end of file
        c.skip(8);                           // driver info
        c.skip(8);                           // root entry: link name offset
        uint64_t rootHeader = c.u(8);
        parseObject(base + rootHeader, "", "", 0);
    } else if (version == 2 || version == 3) {
        int so = (int)c.u(1), sl = (int)c.u(1);
        if (so != 8 || sl != 8) throw Error("only 8-byte offsets/lengths are supported");
        c.skip(1);
        base = c.u(8);
        if (base == kUndef) base = 0;
        c.skip(8); // superblock extension
        c.skip(8); // end of file
        uint64_t rootHeader = c.u(8);
        parseObject(base + rootHeader, "", "", 0);
    } else {
        throw Error("unsupported superblock version " + std::to_string(version));
    }
}

// collect the messages of an object header (v1 or v2), following continuation blocks
static std::vector<Message> headerMessages(const std::vector<uint8_t> &buf, uint64_t addr, uint64_t base) {
    std::vector<Message> out;
    Cursor c(buf, addr);
    c.need(4);
    if (std::memcmp(buf.data() + addr, "OHDR", 4) == 0) { // version 2
        c.skip(4);
        int ver = (int)c.u(1);
        if (ver != 2) throw Error("unsupported object header version " + std::to_string(ver));
        int flags = (int)c.u(1);
        if (flags & 0x20) c.skip(16);
        if (flags & 0x10) c.skip(4);
        uint64_t chunk0 = c.u(1 << (flags & 3));
        struct Chunk { uint64_t at, len; };
        std::vector<Chunk> chunks;
        chunks.push_back({c.p, chunk0});
        for (size_t k = 0; k < chunks.size(); ++k) {
            Cursor m(buf, chunks[k].at);
            uint64_t end = chunks[k].at + chunks[k].len;
            while (m.p + 4 <= end) {
                int type = (int)m.u(1);
                uint64_t size = m.u(2);
                m.skip(1);
                if (flags & 0x04) m.skip(2);
                if (m.p + size > end) break; // gap
                if (type == 0x10) {
                    Cursor cc(buf, m.p);
                    uint64_t off = cc.u(8), len = cc.u(8);
                    // continuation block: "OCHK" + messages + checksum
                    chunks.push_back({base + off + 4, len - 8});
                } else if (type != 0) {
                    out.push_back({type, m.p, size});
                }
                m.skip(size);
            }
        }
        return out;
    }
    int ver = (int)c.u(1);
    if (ver != 1) throw Error("unsupported object header version " + std::to_string(ver) + " at " + std::to_string(addr));
    c.skip(1);
    int nmsg = (int)c.u(2);
    c.skip(4);
    uint64_t hsize = c.u(4);
    c.skip(4); // alignment padding
    struct Chunk { uint64_t at, len; };
    std::vector<Chunk> chunks;
    chunks.push_back({c.p, hsize});
    int seen = 0;
    for (size_t k = 0; k < chunks.size() && seen < nmsg; ++k) {
        Cursor m(buf, chunks[k].at);
        uint64_t end = chunks[k].at + chunks[k].len;
        while (m.p + 8 <= end && seen < nmsg) {
            int type = (int)m.u(2);
            uint64_t size = m.u(2);
            m.skip(4);
            ++seen;
            if (type == 0x10) {
                Cursor cc(buf, m.p);
                uint64_t off = cc.u(8), len = cc.u(8);
                chunks.push_back({base + off, len});
            } else if (type != 0) {
                out.push_back({type, m.p, size});
            }
            m.skip(size);
        }
    }
    return out;
}

void File::parseObject(uint64_t headerAddr, const std::string &name, const std::string &prefix, int depth) {
    if (depth > 8) throw Error("group nesting too deep");
    std::vector<Message> msgs = headerMessages(buf, headerAddr, base);
    DatasetInfo d;
    bool haveSpace = false, haveType = false, haveLayout = false;
    for (const Message &m : msgs) {
        Cursor c(buf, m.at);
        if (m.type == 0x11) { // symbol table message: old-style group
            uint64_t bt = c.u(8), heap = c.u(8);
            parseGroup(base + bt, base + heap, name.empty() ? prefix : prefix + name + "/", depth + 1);
        } else if (m.type == 0x06) { // link message: new-style compact group
            int ver = (int)c.u(1);
            if (ver != 1) throw Error("unsupported link message version");
            int fl = (int)c.u(1);
            int ltype = 0;
            if (fl & 0x08) ltype = (int)c.u(1);
            if (fl & 0x04) c.skip(8);
            if (fl & 0x10) c.skip(1);
            uint64_t nlen = c.u(1 << (fl & 3));
            c.need(nlen);
            std::string lname((const char *)buf.data() + c.p, (size_t)nlen);
            c.skip(nlen);
            if (ltype == 0) {
                uint64_t addr = c.u(8);
                parseObject(base + addr, lname, name.empty() ? prefix : prefix + name + "/", depth + 1);
            }
        } else if (m.type == 0x02) { // link info: dense storage (fractal heap) is not supported
            Cursor li(buf, m.at);
            li.skip(1);
            int fl = (int)li.u(1);
            if (fl & 1) li.skip(8);
            uint64_t fheap = li.u(8);
            if (fheap != kUndef) throw Error("group with dense link storage (more than 8 links, libver=latest) is not supported");
        } else if (m.type == 0x01) {
            int ver = (int)c.u(1);
            int rank = (int)c.u(1);
            int fl = (int)c.u(1);
            if (ver == 1) {
                c.skip(5);
            } else if (ver == 2) {
                c.skip(1);
            } else {
                throw Error("unsupported dataspace version");
            }
            (void)fl;
            d.dims.clear();
            for (int k = 0; k < rank; ++k) d.dims.push_back(c.u(8));
            haveSpace = true;
        } else if (m.type == 0x03) {
            int cv = (int)c.u(1);
            int cls = cv & 0x0f;
            int b0 = (int)c.u(1);
            c.skip(2);
            d.elemSize = (uint32_t)c.u(4);
            d.bigEndian = (b0 & 1) != 0;
            if (cls == 0) {
                d.kind = (b0 & 0x08) ? Kind::Int : Kind::UInt;
            } else if (cls == 1) {
                d.kind = Kind::Float;
                if (d.elemSize != 4 && d.elemSize != 8) throw Error("unsupported float size in '" + name + "'");
            } else {
                d.elemSize = 0; // unsupported class: only an error if the dataset is read
            }
            haveType = true;
        } else if (m.type == 0x08) {
            int ver = (int)c.u(1);
            if (ver == 1 || ver == 2) {
                int dimn = (int)c.u(1);
                int cls = (int)c.u(1);
                c.skip(5);
                if (cls == 1) {
                    uint64_t a = c.u(8);
                    d.address = a == kUndef ? kUndef : base + a;
                    uint64_t bytes = 1;
                    for (int k = 0; k < dimn; ++k) bytes *= c.u(4);
                    d.byteSize = bytes;
                    haveLayout = true;
                } else if (cls == 0) {
                    for (int k = 0; k < dimn; ++k) c.skip(4);
                    d.byteSize = c.u(4);
                    d.address = c.p;
                    haveLayout = true;
                }
            } else if (ver == 3 || ver == 4) {
                int cls = (int)c.u(1);
                if (cls == 1) {
                    uint64_t a = c.u(8);
                    d.address = a == kUndef ? kUndef : base + a;
                    d.byteSize = c.u(8);
                    haveLayout = true;
                } else if (cls == 0) {
                    d.byteSize = c.u(2);
                    d.address = c.p;
                    haveLayout = true;
                }
            }
        }
    }
    if (haveSpace && haveType) {
        d.name = prefix + name;
        if (!haveLayout) d.address = kUndef; // chunked / virtual: listed, but reading it is an error
        sets[d.name] = d;
    }
}

void File::parseGroup(uint64_t btreeAddr, uint64_t heapAddr, const std::string &prefix, int depth) {
    Cursor h(buf, heapAddr);
    h.need(32);
    if (std::memcmp(buf.data() + heapAddr, "HEAP", 4) != 0) throw Error("bad local heap signature");
    h.skip(8);
    h.skip(8); // data segment size
    h.skip(8); // free list head
    uint64_t heapData = base + h.u(8);
    parseBtree(btreeAddr, heapData, prefix, depth);
}

void File::parseBtree(uint64_t addr, uint64_t heapData, const std::string &prefix, int depth) {
    Cursor c(buf, addr);
    c.need(24);
    if (std::memcmp(buf.data() + addr, "TREE", 4) != 0) throw Error("bad B-tree signature");
    c.skip(4);
    int type = (int)c.u(1), level = (int)c.u(1);
    int used = (int)c.u(2);
    if (type != 0) throw Error("unexpected B-tree node type");
    c.skip(16); // siblings
    for (int k = 0; k < used; ++k) {
        c.skip(8); // key k
        uint64_t child = base + c.u(8);
        if (level > 0) {
            parseBtree(child, heapData, prefix, depth);
            continue;
        }
        Cursor s(buf, child);
        s.need(8);
        if (std::memcmp(buf.data() + child, "SNOD", 4) != 0) throw Error("bad symbol table node signature");
        s.skip(6);
        int nsym = (int)s.u(2);
        for (int e = 0; e < nsym; ++e) {
            uint64_t nameOff = s.u(8), header = s.u(8);
            s.skip(24);
            uint64_t np = heapData + nameOff;
            if (np >= buf.size()) throw Error("link name outside the file");
            std::string nm((const char *)buf.data() + np, strnlen((const char *)buf.data() + np, buf.size() - np));
            parseObject(base + header, nm, prefix, depth);
        }
    }
}

bool File::exist(const std::string &name) const { return sets.count(normName(name)) != 0; }

const DatasetInfo &File::info(const std::string &name) const {
    auto it = sets.find(normName(name));
    if (it == sets.end()) throw Error("no dataset '" + name + "'");
    return it->second;
}

std::vector<std::string> File::listDatasets() const {
    std::vector<std::string> r;
    for (const auto &kv : sets) r.push_back(kv.first);
    return r;
}

template <typename T> void File::readConverted(const DatasetInfo &d, std::vector<T> &out) const {
    const uint64_t n = d.numElements();
    if (d.address == kUndef) {
        if (n == 0) {
            out.clear();
            return;
        }
        throw Error("dataset '" + d.name + "' has no contiguous storage (chunked/compressed layouts are not supported)");
    }
    if (d.elemSize == 0 || d.elemSize > 8) throw Error("dataset '" + d.name + "' has an unsupported datatype");
    if (d.address + n * d.elemSize > buf.size()) throw Error("dataset '" + d.name + "' extends past the end of the file");
    out.resize((size_t)n);
    const uint8_t *src = buf.data() + d.address;
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t raw[8] = {0};
        for (uint32_t k = 0; k < d.elemSize; ++k) raw[k] = src[i * d.elemSize + (d.bigEndian ? d.elemSize - 1 - k : k)];
        if (d.kind == Kind::Float) {
            if (d.elemSize == 8) {
                double v;
                std::memcpy(&v, raw, 8);
                out[i] = (T)v;
            } else {
                float v;
                std::memcpy(&v, raw, 4);
                out[i] = (T)v;
            }
        } else {
            uint64_t v = 0;
            std::memcpy(&v, raw, 8);
            if (d.kind == Kind::Int && d.elemSize < 8 && (v >> (8 * d.elemSize - 1)) & 1) v |= ~0ull << (8 * d.elemSize);
            out[i] = d.kind == Kind::Int ? (T)(int64_t)v : (T)v;
        }
    }
}

void File::read(const std::string &name, std::vector<double> &out) const { readConverted(info(name), out); }
void File::read(const std::string &name, std::vector<int> &out) const { readConverted(info(name), out); }
void File::read(const std::string &name, std::vector<std::vector<double>> &out) const {
    const DatasetInfo &d = info(name);
    if (d.dims.size() != 2) throw Error("dataset '" + d.name + "' is not two-dimensional");
    std::vector<double> flat;
    readConverted(d, flat);
    out.assign((size_t)d.dims[0], std::vector<double>((size_t)d.dims[1]));
    for (uint64_t i = 0; i < d.dims[0]; ++i)
        for (uint64_t k = 0; k < d.dims[1]; ++k) out[i][k] = flat[i * d.dims[1] + k];
}

// ------------------------------------------------------------------------------------------------
// writer
// ------------------------------------------------------------------------------------------------
namespace {

struct Out {
    std::vector<uint8_t> b;
    void u(uint64_t v, int n) {
        for (int k = 0; k < n; ++k) b.push_back((uint8_t)(v >> (8 * k)));
    }
    void bytes(const void *p, size_t n) {
        const uint8_t *q = (const uint8_t *)p;
        b.insert(b.end(), q, q + n);
    }
    void pad8() {
        while (b.size() % 8) b.push_back(0);
    }
    void zeros(size_t n) { b.insert(b.end(), n, 0); }
    void patch(size_t at, uint64_t v, int n) {
        for (int k = 0; k < n; ++k) b[at + k] = (uint8_t)(v >> (8 * k));
    }
};

const int kLeafK = 4, kInternalK = 16; // libhdf5 defaults
const size_t kSnodSize = 8 + 2 * kLeafK * 40;
const size_t kBtreeSize = 24 + (2 * kInternalK + 1) * 8 + 2 * kInternalK * 8;

void messageHeader(Out &o, int type, size_t size, int flags) {
    o.u(type, 2);
    o.u(size, 2);
    o.u(flags, 1);
    o.u(0, 3);
}

} // namespace

Writer::Writer(const std::string &p) : path(p) {}

Writer::~Writer() {
    if (!closed) {
        try {
            close();
        } catch (...) {
        }
    }
}

void Writer::add(const std::string &name, const std::vector<uint64_t> &dims, int type, const void *data, size_t elem) {
    if (closed) throw Error("write after close");
    Pending p;
    p.name = normName(name);
    if (p.name.empty() || p.name.find('/') != std::string::npos) throw Error("datasets live in the root group: '" + name + "'");
    for (const Pending &q : items)
        if (q.name == p.name) throw Error("dataset '" + name + "' written twice");
    p.dims = dims.empty() ? std::vector<uint64_t>{1} : dims;
    p.type = type;
    uint64_t n = 1;
    for (uint64_t d : p.dims) n *= d;
    p.bytes.resize((size_t)(n * elem));
    if (n) std::memcpy(p.bytes.data(), data, p.bytes.size());
    items.push_back(std::move(p));
}
void Writer::write(const std::string &n, const std::vector<uint64_t> &d, const double *p) { add(n, d, 0, p, 8); }
void Writer::write(const std::string &n, const std::vector<uint64_t> &d, const int32_t *p) { add(n, d, 1, p, 4); }
void Writer::write(const std::string &n, const std::vector<uint64_t> &d, const int8_t *p) { add(n, d, 2, p, 1); }

void Writer::close() {
    if (closed) return;
    closed = true;
    std::sort(items.begin(), items.end(), [](const Pending &a, const Pending &b) { return std::strcmp(a.name.c_str(), b.name.c_str()) < 0; });
    const size_t nset = items.size();
    const size_t nsnod = std::max<size_t>(1, (nset + 2 * kLeafK - 1) / (2 * kLeafK));
    if (nsnod > (size_t)2 * kInternalK) throw Error("too many datasets for one B-tree node");

    // ---- local heap data segment: "" at 0, names 8-aligned, one free block at the end ----
    std::vector<uint64_t> nameOff(nset);
    Out heap;
    heap.zeros(8);
    for (size_t k = 0; k < nset; ++k) {
        nameOff[k] = heap.b.size();
        heap.bytes(items[k].name.c_str(), items[k].name.size() + 1);
        heap.pad8();
    }
    const uint64_t freeOff = heap.b.size();
    heap.u(1, 8);  // next free block: H5HL_FREE_NULL
    heap.u(32, 8); // size of this free block
    heap.zeros(16);

    // ---- layout ----
    const uint64_t atSuper = 0, atRootHdr = 96, atHeapHdr = atRootHdr + 16 + 24, atHeapData = atHeapHdr + 32;
    const uint64_t atBtree = atHeapData + heap.b.size();
    const uint64_t atSnod = atBtree + kBtreeSize;
    uint64_t at = atSnod + nsnod * kSnodSize;
    std::vector<uint64_t> hdrAt(nset), dataAt(nset);
    std::vector<Out> hdr(nset);
    for (size_t k = 0; k < nset; ++k) {
        const Pending &p = items[k];
        Out &o = hdr[k];
        // messages first (their total size goes into the prefix)
        Out m;
        const size_t rank = p.dims.size();
        messageHeader(m, 0x0001, 8 + 8 * rank, 0); // dataspace, version 1
        m.u(1, 1);
        m.u(rank, 1);
        m.u(0, 1);
        m.u(0, 5);
        for (uint64_t d : p.dims) m.u(d, 8);
        if (p.type == 0) { // IEEE f64 little-endian
            messageHeader(m, 0x0003, 24, 1);
            m.u(0x11, 1);
            m.u(0x20, 1);
            m.u(0x3f, 1);
            m.u(0, 1);
            m.u(8, 4);
            m.u(0, 2);
            m.u(64, 2);
            m.u(52, 1);
            m.u(11, 1);
            m.u(0, 1);
            m.u(52, 1);
            m.u(1023, 4);
            m.u(0, 4);
        } else { // signed two's-complement little-endian integer
            const int size = p.type == 1 ? 4 : 1;
            messageHeader(m, 0x0003, 16, 1);
            m.u(0x10, 1);
            m.u(0x08, 1);
            m.u(0, 2);
            m.u(size, 4);
            m.u(0, 2);
            m.u(8 * size, 2);
            m.u(0, 4);
        }
        messageHeader(m, 0x0005, 8, 1); // fill value, version 2: early allocation, fill time "if set", undefined
        m.u(2, 1);
        m.u(1, 1);
        m.u(2, 1);
        m.u(0, 1);
        m.u(0, 4);
        messageHeader(m, 0x0008, 24, 0); // layout, version 3, contiguous
        m.u(3, 1);
        m.u(1, 1);
        const size_t addrPatch = m.b.size();
        m.u(0, 8);
        m.u(p.bytes.size(), 8);
        m.u(0, 6);
        o.u(1, 1); // object header version 1
        o.u(0, 1);
        o.u(4, 2); // messages
        o.u(1, 4); // reference count
        o.u(m.b.size(), 4);
        o.u(0, 4);
        hdrAt[k] = at;
        dataAt[k] = at + 16 + m.b.size();
        dataAt[k] = (dataAt[k] + 7) / 8 * 8;
        m.patch(addrPatch, p.bytes.empty() ? kUndef : dataAt[k], 8);
        o.bytes(m.b.data(), m.b.size());
        at = dataAt[k] + p.bytes.size();
        at = (at + 7) / 8 * 8;
    }
    const uint64_t eof = at;

    Out f;
    // ---- superblock, version 0 ----
    f.bytes(kSignature, 8);
    f.u(0, 1); // superblock version
    f.u(0, 1); // free-space storage version
    f.u(0, 1); // root group symbol table entry version
    f.u(0, 1);
    f.u(0, 1); // shared header message format version
    f.u(8, 1); // size of offsets
    f.u(8, 1); // size of lengths
    f.u(0, 1);
    f.u(kLeafK, 2);
    f.u(kInternalK, 2);
    f.u(0, 4);      // file consistency flags
    f.u(0, 8);      // base address
    f.u(kUndef, 8); // free-space info
    f.u(eof, 8);    // end of file
    f.u(kUndef, 8); // driver information block
    f.u(0, 8);      // root entry: link name offset
    f.u(atRootHdr, 8);
    f.u(1, 4); // cache type 1: B-tree and heap addresses in the scratch pad
    f.u(0, 4);
    f.u(atBtree, 8);
    f.u(atHeapHdr, 8);
    (void)atSuper;
    // ---- root group object header: one symbol table message ----
    f.u(1, 1);
    f.u(0, 1);
    f.u(1, 2);
    f.u(1, 4);
    f.u(24, 4);
    f.u(0, 4);
    messageHeader(f, 0x0011, 16, 0);
    f.u(atBtree, 8);
    f.u(atHeapHdr, 8);
    // ---- local heap ----
    f.bytes("HEAP", 4);
    f.u(0, 4);
    f.u(heap.b.size(), 8);
    f.u(freeOff, 8);
    f.u(atHeapData, 8);
    f.bytes(heap.b.data(), heap.b.size());
    // ---- B-tree node (group node, leaf level) ----
    {
        Out t;
        t.bytes("TREE", 4);
        t.u(0, 1);
        t.u(0, 1);
        t.u(nset ? nsnod : 0, 2);
        t.u(kUndef, 8);
        t.u(kUndef, 8);
        t.u(0, 8); // key 0: the empty name
        for (size_t s = 0; s < nsnod && nset; ++s) {
            t.u(atSnod + s * kSnodSize, 8);
            size_t last = std::min(nset, (s + 1) * 2 * kLeafK) - 1;
            t.u(nameOff[last], 8);
        }
        t.zeros(kBtreeSize - t.b.size());
        f.bytes(t.b.data(), t.b.size());
    }
    // ---- symbol table nodes ----
    for (size_t s = 0; s < nsnod; ++s) {
        Out n;
        size_t first = s * 2 * kLeafK, last = std::min(nset, first + 2 * kLeafK);
        n.bytes("SNOD", 4);
        n.u(1, 1);
        n.u(0, 1);
        n.u(last > first ? last - first : 0, 2);
        for (size_t k = first; k < last; ++k) {
            n.u(nameOff[k], 8);
            n.u(hdrAt[k], 8);
            n.u(0, 4);
            n.u(0, 4);
            n.zeros(16);
        }
        n.zeros(kSnodSize - n.b.size());
        f.bytes(n.b.data(), n.b.size());
    }
    // ---- datasets ----
    for (size_t k = 0; k < nset; ++k) {
        if (f.b.size() != hdrAt[k]) throw Error("internal layout error");
        f.bytes(hdr[k].b.data(), hdr[k].b.size());
        f.zeros((size_t)(dataAt[k] - f.b.size()));
        f.bytes(items[k].bytes.data(), items[k].bytes.size());
        f.pad8();
    }
    if (f.b.size() != eof) throw Error("internal layout error (eof)");
    FILE *fp = std::fopen(path.c_str(), "wb");
    if (!fp) throw Error("cannot create '" + path + "'");
    size_t w = std::fwrite(f.b.data(), 1, f.b.size(), fp);
    if (std::fclose(fp) != 0 || w != f.b.size()) throw Error("write error on '" + path + "'");
    items.clear();
}

} // namespace H5Lite

// ------------------------------------------------------------------------------------------------
// C entry points (ctypes: meshlesshydro_b200/h5lite.py uses them to make IC files and read snapshots)
// ------------------------------------------------------------------------------------------------
extern "C" {

static thread_local std::string g_h5err;
const char *h5lite_last_error() { return g_h5err.c_str(); }

void *h5lite_open(const char *path) {
    try {
        return new H5Lite::File(path);
    } catch (const std::exception &e) {
        g_h5err = e.what();
        return nullptr;
    }
}
void h5lite_close(void *f) { delete (H5Lite::File *)f; }
int h5lite_num_datasets(void *f) { return (int)((H5Lite::File *)f)->listDatasets().size(); }
int h5lite_dataset_name(void *f, int k, char *out, int cap) {
    auto names = ((H5Lite::File *)f)->listDatasets();
    if (k < 0 || k >= (int)names.size()) return -1;
    std::snprintf(out, cap, "%s", names[k].c_str());
    return 0;
}
// rank, dims[<=8], kind (0 float, 1 int, 2 uint), element size; returns 0 or -1
int h5lite_info(void *f, const char *name, int *rank, unsigned long long *dims, int *kind, int *elem) {
    try {
        const H5Lite::DatasetInfo &d = ((H5Lite::File *)f)->info(name);
        *rank = (int)d.dims.size();
        for (size_t k = 0; k < d.dims.size() && k < 8; ++k) dims[k] = d.dims[k];
        *kind = d.kind == H5Lite::Kind::Float ? 0 : (d.kind == H5Lite::Kind::Int ? 1 : 2);
        *elem = (int)d.elemSize;
        return 0;
    } catch (const std::exception &e) {
        g_h5err = e.what();
        return -1;
    }
}
long long h5lite_read_f64(void *f, const char *name, double *out, long long cap) {
    try {
        std::vector<double> v;
        ((H5Lite::File *)f)->read(name, v);
        if ((long long)v.size() > cap) return -2;
        std::memcpy(out, v.data(), v.size() * sizeof(double));
        return (long long)v.size();
    } catch (const std::exception &e) {
        g_h5err = e.what();
        return -1;
    }
}
long long h5lite_read_i32(void *f, const char *name, int *out, long long cap) {
    try {
        std::vector<int> v;
        ((H5Lite::File *)f)->read(name, v);
        if ((long long)v.size() > cap) return -2;
        std::memcpy(out, v.data(), v.size() * sizeof(int));
        return (long long)v.size();
    } catch (const std::exception &e) {
        g_h5err = e.what();
        return -1;
    }
}
void *h5lite_create(const char *path) { return new H5Lite::Writer(path); }
// type: 0 f64, 1 i32, 2 i8
int h5lite_write(void *w, const char *name, int rank, const unsigned long long *dims, int type, const void *data) {
    try {
        std::vector<uint64_t> d(dims, dims + rank);
        H5Lite::Writer *wr = (H5Lite::Writer *)w;
        if (type == 0) wr->write(name, d, (const double *)data);
        else if (type == 1) wr->write(name, d, (const int32_t *)data);
        else if (type == 2) wr->write(name, d, (const int8_t *)data);
        else return -1;
        return 0;
    } catch (const std::exception &e) {
        g_h5err = e.what();
        return -1;
    }
}
int h5lite_finish(void *w) {
    int rc = 0;
    try {
        ((H5Lite::Writer *)w)->close();
    } catch (const std::exception &e) {
        g_h5err = e.what();
        rc = -1;
    }
    delete (H5Lite::Writer *)w;
    return rc;
}

} // extern "C"
