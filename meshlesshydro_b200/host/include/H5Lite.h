// H5Lite -- minimal self-contained HDF5 reader/writer for the two file kinds of the demonstrator.
//
// The reference reads initial conditions and writes snapshots through HighFive/libhdf5
// (demonstrator/src/InitialDistribution.cpp:7-30, demonstrator/src/Particles.cpp:2978-3076); neither is
// available in this image, so this file implements the subset of the HDF5 File Format Specification
// (version 1.1 structures: superblock version 0, version-1 object headers, symbol-table groups with a
// version-1 B-tree + local heap, contiguous little-endian datasets) that
//   * h5py / libhdf5 write by default for small flat files  -> H5Lite::File::open can read them, and
//   * libhdf5 (HighFive, h5py, h5dump) reads               -> files written by H5Lite::Writer.
// Datasets: 1-D or 2-D, IEEE f32/f64 and signed/unsigned 8/16/32/64-bit integers (read, converted to double or
// int), f64 / i32 / i8 (write).  Datasets live in the root group.  No chunking, compression, attributes.
#ifndef MLH_H5LITE_H
#define MLH_H5LITE_H

#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace H5Lite {

struct Error : std::runtime_error {
    explicit Error(const std::string &m) : std::runtime_error("H5Lite: " + m) {}
};

enum class Kind { Float, Int, UInt };

struct DatasetInfo {
    std::string name;
    std::vector<uint64_t> dims; // rank 0 (scalar) .. n
    Kind kind = Kind::Float;
    uint32_t elemSize = 8;
    bool bigEndian = false;
    uint64_t address = 0;  // absolute file offset of the contiguous raw data
    uint64_t byteSize = 0;
    uint64_t numElements() const {
        uint64_t n = 1;
        for (uint64_t d : dims) n *= d;
        return n;
    }
};

// ---- reader ----
class File {
public:
    explicit File(const std::string &path); // throws H5Lite::Error
    bool exist(const std::string &name) const;
    const DatasetInfo &info(const std::string &name) const; // name with or without leading '/'
    std::vector<std::string> listDatasets() const;
    // element-wise converted reads (row-major)
    void read(const std::string &name, std::vector<double> &out) const;
    void read(const std::string &name, std::vector<int> &out) const;
    // 2-D dataset as vector of rows (what HighFive's DataSet::read(std::vector<std::vector<double>>&) gives)
    void read(const std::string &name, std::vector<std::vector<double>> &out) const;

private:
    std::vector<uint8_t> buf; // whole file (the files of this path are particle tables; read once)
    uint64_t base = 0;        // superblock base address (user block)
    std::map<std::string, DatasetInfo> sets;
    void parseGroup(uint64_t btreeAddr, uint64_t heapAddr, const std::string &prefix, int depth);
    void parseBtree(uint64_t addr, uint64_t heapData, const std::string &prefix, int depth);
    void parseObject(uint64_t headerAddr, const std::string &name, const std::string &prefix, int depth);
    template <typename T> void readConverted(const DatasetInfo &d, std::vector<T> &out) const;
};

// ---- writer ----
class Writer {
public:
    explicit Writer(const std::string &path); // truncates; nothing is written before close()
    ~Writer();
    void write(const std::string &name, const std::vector<uint64_t> &dims, const double *data);
    void write(const std::string &name, const std::vector<uint64_t> &dims, const int32_t *data);
    void write(const std::string &name, const std::vector<uint64_t> &dims, const int8_t *data);
    void close(); // lays out and writes the file; throws on I/O errors

private:
    struct Pending {
        std::string name;
        std::vector<uint64_t> dims;
        int type; // 0 f64, 1 i32, 2 i8
        std::vector<uint8_t> bytes;
    };
    std::string path;
    std::vector<Pending> items;
    bool closed = false;
    void add(const std::string &name, const std::vector<uint64_t> &dims, int type, const void *data, size_t elem);
};

} // namespace H5Lite

#endif
