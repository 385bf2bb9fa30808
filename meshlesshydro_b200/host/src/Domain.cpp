#include "../include/Domain.h"

Domain::Domain(Cell b) : bounds{b} {}
Domain::~Domain() {}

void Domain::createGrid(const double &kernelSize) {
    // cells per axis = floor(extent / h), so that a cell edge is >= h (Domain.cpp:10-22)
    cellsX = (int)floor((bounds.maxX - bounds.minX) / kernelSize);
    cellsY = (int)floor((bounds.maxY - bounds.minY) / kernelSize);
    long total = (long)cellsX * cellsY;
    cellSizeX = (bounds.maxX - bounds.minX) / (double)cellsX;
    cellSizeY = (bounds.maxY - bounds.minY) / (double)cellsY;
    Logger(DEBUG) << "      > cellSizeX = " << cellSizeX << ", cellsX = " << cellsX;
    Logger(DEBUG) << "      > cellSizeY = " << cellSizeY << ", cellsY = " << cellsY;
#if DIM == 3
    cellsZ = (int)floor((bounds.maxZ - bounds.minZ) / kernelSize);
    total *= cellsZ;
    cellSizeZ = (bounds.maxZ - bounds.minZ) / (double)cellsZ;
    Logger(DEBUG) << "      > cellSizeZ = " << cellSizeZ << ", cellsZ = " << cellsZ;
#endif
    numGridCells = (int)total;
    grid.clear();
    if (total <= 0 || total > MATERIALIZE_LIMIT) return;
    grid.resize((size_t)total);
#if DIM == 2
    for (int iY = 0; iY < cellsY; ++iY)
        for (int iX = 0; iX < cellsX; ++iX) {
            Cell &c = grid[iX + iY * cellsX];
            c.minX = iX * cellSizeX + bounds.minX;
            c.maxX = (iX + 1) * cellSizeX + bounds.minX;
            c.minY = iY * cellSizeY + bounds.minY;
            c.maxY = (iY + 1) * cellSizeY + bounds.minY;
        }
#else
    for (int iZ = 0; iZ < cellsZ; ++iZ)
        for (int iY = 0; iY < cellsY; ++iY)
            for (int iX = 0; iX < cellsX; ++iX) {
                Cell &c = grid[iX + iY * cellsX + (size_t)iZ * cellsX * cellsY];
                c.minX = iX * cellSizeX + bounds.minX;
                c.maxX = (iX + 1) * cellSizeX + bounds.minX;
                c.minY = iY * cellSizeY + bounds.minY;
                c.maxY = (iY + 1) * cellSizeY + bounds.minY;
                c.minZ = iZ * cellSizeZ + bounds.minZ;
                c.maxZ = (iZ + 1) * cellSizeZ + bounds.minZ;
            }
#endif
}

// 3^DIM stencil in the reference's order: x outer, then y, then z; -1 outside the grid (no periodic wrap,
// Domain.cpp:83-118).  The device search uses the same order (csrc/k2_neighbours.cu).
void Domain::getNeighborCells(const int &iCell, int *neighborCell) {
    const int iX = iCell % cellsX;
    const int iY = (iCell / cellsX) % cellsY;
    int n = 0;
#if DIM == 2
    for (int k = iX - 1; k <= iX + 1; ++k)
        for (int l = iY - 1; l <= iY + 1; ++l)
            neighborCell[n++] = (k < 0 || k >= cellsX || l < 0 || l >= cellsY) ? -1 : k + l * cellsX;
#else
    const int iZ = iCell / (cellsX * cellsY);
    for (int k = iX - 1; k <= iX + 1; ++k)
        for (int l = iY - 1; l <= iY + 1; ++l)
            for (int m = iZ - 1; m <= iZ + 1; ++m)
                neighborCell[n++] = (k < 0 || k >= cellsX || l < 0 || l >= cellsY || m < 0 || m >= cellsZ)
                                        ? -1
                                        : k + l * cellsX + m * cellsX * cellsY;
#endif
}

void Domain::printout() {
    Logger(INFO) << "Domain > X [" << bounds.minX << ", " << bounds.maxX << "]";
    Logger(INFO) << "         Y [" << bounds.minY << ", " << bounds.maxY << "]";
#if DIM == 3
    Logger(INFO) << "         Z [" << bounds.minZ << ", " << bounds.maxZ << "]";
#endif
}
