// Search-grid description with the public surface of the reference's Domain
// (/root/reference/demonstrator/include/Domain.h:14-75).  In the GPU build the per-cell particle lists live in
// HBM as contiguous ranges of the cell-sorted particle arrays (csrc/k1_sort.cu), so Cell::prtcls of the grid
// cells stays empty; cell counts, cell sizes and bounds are the values the device grid uses (same double
// arithmetic as Domain::createGrid, Domain.cpp:9-54).
#ifndef DEMONSTRATOR_DOMAIN_H
#define DEMONSTRATOR_DOMAIN_H

#include <cmath>
#include <vector>

#include "parameter.h"
#include "Logger.h"

class Domain {
public:
    struct Cell {
        Cell(double *bounds)
            : minX{bounds[0]}, maxX{bounds[DIM]}, minY{bounds[1]}, maxY{bounds[DIM + 1]}
#if DIM == 3
              , minZ{bounds[2]}, maxZ{bounds[DIM + 2]}
#endif
        {}
        Cell() {}
        double minX{0.}, maxX{0.}, minY{0.}, maxY{0.};
#if DIM == 3
        double minZ{0.}, maxZ{0.};
#endif
        std::vector<int> prtcls{};
    };

    Domain(Cell bounds);
    ~Domain();

    int numGridCells{0};
    std::vector<Cell> grid; // cell bounds; filled up to MATERIALIZE_LIMIT cells (the device never reads it)
    static constexpr long MATERIALIZE_LIMIT = 1L << 24;

    void createGrid(const double &kernelSize);
    void getNeighborCells(const int &iCell, int *neighborCells);
    void printout();

    Cell bounds; // global cell
    int cellsX{0}, cellsY{0};
#if DIM == 3
    int cellsZ{0};
#endif
    double cellSizeX{0}, cellSizeY{0};
#if DIM == 3
    double cellSizeZ{0};
#endif
};

#endif // DEMONSTRATOR_DOMAIN_H
