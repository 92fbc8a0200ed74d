// Smoothed-aggregation multigrid hierarchy for the Krylov inner solves (host setup, amg_setup.cpp; device cycle, amg.cu).
// The hierarchy is built ONCE per problem from the real operator slots: every level carries the Galerkin coarse
// operator of EACH slot on one shared pattern, so the shifted operator of a contour node, sum_i c_i slot_i, is
// assembled per level with the node's coefficients exactly like the fine one (A_c - z B_c = P^T (A - z B) P).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

struct AmgHostLevel {
    int n = 0;
    std::vector<int> rowptr, col;             // CSR pattern shared by the slots (level 0: the natural union pattern)
    std::vector<std::vector<double>> vals;    // [nslots][nnz]
    std::vector<int> dpos;                    // position of the diagonal entry of each row
    double rho = 0.0;                         // estimate of the spectral radius of D^-1 A (slot 0)
    int nc = 0;                               // size of the next level (0: this is the coarsest one)
    std::vector<int> p_rowptr, p_col;         // prolongator P (n x nc), smoothed aggregation
    std::vector<double> p_val;
    std::vector<int> r_rowptr, r_col;         // restriction R = P^T (nc x n)
    std::vector<double> r_val;
};

struct AmgHost {
    std::vector<AmgHostLevel> levels;
    bool ok = false;
    std::string why;                          // reason when no usable hierarchy could be built
    double setup_seconds = 0.0;
};

// rowptr/col/vals: natural-order union pattern with sorted rows; slot 0 drives aggregation and prolongator smoothing.
// Coarsening stops at max_coarse rows (the coarsest level is solved by dense LU on the device).
void amg_setup_host(int64_t n, const int64_t* rowptr, const int* col, int nslots, const double* const* vals, int max_coarse,
                    AmgHost& out);
// P^T of a CSR matrix with nrows x ncols
void amg_transpose(int nrows, int ncols, const std::vector<int>& rp, const std::vector<int>& ci, const std::vector<double>& v,
                   std::vector<int>& trp, std::vector<int>& tci, std::vector<double>& tv);
int host_threads();
