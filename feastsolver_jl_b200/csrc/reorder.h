// Tile plan of the sparse operators (reorder.cpp): row ordering, tile boundaries, halo lists.
#pragma once
#include <stdint.h>

#include <vector>

struct TileCaps {
    int rows_cap;   // rows of the input block a CTA can hold in shared memory (tile rows + halo rows)
    int64_t nnz_cap; // (padded) nonzeros of a tile whose CSR slice is staged in shared memory
    int tile_max;   // max rows of a tile
    int domain_rows; // > 0: tiles are grown inside compact domains of about this many rows (L2 locality of the sweep)
};

struct TilePlan {
    std::vector<int> order;      // new -> old row map (identity when the rows are not renumbered)
    std::vector<int> tile_ptr;   // [ntiles + 1] first row of each tile (new numbering)
    std::vector<int> halo_ptr;   // [ntiles + 1]
    std::vector<int> halo_idx;   // rows (new numbering) referenced by a tile from outside, sorted per tile
    bool ok = true;              // every tile fits the capacities
    double halo_ratio = 0.0;     // halo rows per row
};

// ordering + tile boundaries from the pattern in the caller's numbering
void build_tile_order(int64_t n, const int64_t* rowptr, const int* col, bool reorder, const TileCaps& caps, TilePlan& plan);
// halo lists and 16-bit tile-local column numbers from the pattern in the NEW numbering; returns nonzero on overflow
int build_tile_halo(int64_t n, const int64_t* rowptr, const int* col, TilePlan& plan, std::vector<uint16_t>& lcol);
