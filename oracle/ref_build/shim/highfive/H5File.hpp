/*
 * oracle/ref_build/shim/highfive/H5File.hpp -- TEST INFRASTRUCTURE.
 *
 * No-op stand-in for HighFive (absent from this image; HDF5 itself is absent too).
 * Only Particles::dump2file / dumpNNL (Particles.cpp:2695-2789, 2978-3076) and
 * InitialDistribution (not compiled into oracle/_ref) touch it; the MFV arithmetic
 * does not.  Writes are discarded.
 */
#ifndef MLH_SHIM_HIGHFIVE_H5FILE_HPP
#define MLH_SHIM_HIGHFIVE_H5FILE_HPP

#include <sstream>
#include <string>
#include <vector>
#include <cstddef>

namespace HighFive {

class DataSpace {
public:
    DataSpace(std::size_t) {}
    DataSpace(const std::vector<std::size_t> &) {}
};

class DataSet {
public:
    template <typename T> void write(const T &) {}
    template <typename T> void read(T &) {}
};

class File {
public:
    enum : unsigned { ReadOnly = 0x00u, ReadWrite = 0x01u, Truncate = 0x02u, Excl = 0x04u, Create = 0x10u };
    File(const std::string &, unsigned = ReadOnly) {}
    template <typename T> DataSet createDataSet(const std::string &, const DataSpace &) { return DataSet(); }
    DataSet getDataSet(const std::string &) { return DataSet(); }
};

} // namespace HighFive

#endif
