// Host-side, once-per-model topology preprocessing.
//
// Replaces what the reference recomputes every step: LHS.__init__ index sets
// (reference linalg.py:28-32), the per-step COO row/col lists (linalg.py:69-156) and the
// COO->CSR conversion with duplicate summation (transport.py:215-218).  Because the sparsity
// pattern of A(t) is time-invariant (n + 2*E_int entries), it is built once:
//   * real cells are renumbered by reverse Cuthill-McKee (optional),
//   * edges are renumbered: internal edges first, sorted by their lower cell; ghost edges after,
//     grouped by cell in ascending ORIGINAL edge id (the order the reference's last-write-wins
//     ghost scatter depends on, linalg.py:349-351, 377-378),
//   * off-diagonal CSR (rowptr, col) with a slot->(edge, side) map: every slot is owned by exactly
//     one (edge, side) pair, so assembly writes each value once with no atomics,
//   * the list of boundary cells with their ghost edges.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace cwr {

constexpr int32_t kLaterBit = (int32_t)0x80000000;
constexpr int32_t kPrevBit = 0x40000000;      // the neighbour has the colour swept immediately before the row's
constexpr int32_t kColMask = 0x3fffffff;

struct Topology {
    int n = 0;        // real cells = matrix order (reference: nreal + 1)
    int F = 0;        // real + ghost cells
    int E = 0;        // edges
    int E_int = 0;    // internal edges (both cells real)
    int E_g = 0;      // ghost edges
    int G = 0;        // ghost cells = F - n
    int64_t nnz = 0;  // off-diagonal entries = 2 * E_int

    std::vector<int32_t> new_of_old;  // (n)  device cell id of reference cell
    std::vector<int32_t> old_of_new;  // (n)
    std::vector<int32_t> eperm;       // (E)  reference edge id of device edge e'
    std::vector<int32_t> f1p, f2p;    // (E)  device cell ids of device edge e'; ghost: n + (f2 - n)

    std::vector<int32_t> rowptr;      // (n+1) off-diagonal CSR
    std::vector<int32_t> col;         // (nnz)
    std::vector<int32_t> slot_edge;   // (nnz) (e' << 1) | side ; side 0: row == f1 of the edge, 1: row == f2

    std::vector<int32_t> bcell;       // (nb)   device ids of cells that own ghost edges
    std::vector<int32_t> bptr;        // (nb+1) ranges into bedge
    std::vector<int32_t> bedge;       // (E_g)  device edge ids, ascending original id within a cell
    int W = 4;                        // ELL width: max row length rounded up to a multiple of 4
    std::vector<int32_t> ell_col;     // (n*W) row-major; padding points at the row itself; bit 31 (kLaterBit): the
                                      // neighbour's colour is >= the row's (visited later in a Gauss-Seidel sweep);
                                      // bit 30 (kPrevBit): the neighbour's colour is the one swept just before the row's
                                      // (the only values a sweep step has to wait for: cwr_kernels.cuh, k_gs_strip)
    std::vector<int32_t> ell_code;    // (n*W) slot_edge code, -1 for padding
    // Gauss-Seidel colours: rows are ordered (part, colour, level, RCM position)
    int n_colors = 0;                 // 0: no colouring (rows in RCM order, parts = equal chunks of it)
    std::vector<int32_t> color_ptr;   // (n_parts, n_colors+1) absolute row ranges of part p's colours
    std::vector<uint8_t> color_of;    // (n) colour of a row (new numbering)
    int n_levels = 1;                 // downstream levels of the flow hint the colours were cut from (diagnostic)
    // domain decomposition (n_parts = 1: everything is one part)
    std::vector<int32_t> part_ptr;    // (n_parts+1) row range owned by part p
    std::vector<uint8_t> send_mask;   // (n) bit q set: part q (not the owner) reads this row
    std::vector<int32_t> send_ptr;    // (n_parts+1) ranges into send_rows
    std::vector<int32_t> send_rows;   // rows with a non-zero send_mask, by owner, ascending
    // strips (neighbour-synchronised Gauss-Seidel sweeps, n_strips > 0): every part is cut into n_strips equal chunks
    // of the RCM order and the rows are ordered (part, strip, colour, RCM position) -- a strip is a contiguous row
    // range owned by ONE CTA of the sweep kernel, which then only waits for the strips its rows are coupled to
    // (color_ptr is not meaningful in this mode)
    int n_strips = 0;                      // strips per part
    std::vector<int32_t> strip_cptr;       // (n_parts * n_strips, n_colors+1) absolute row ranges of a strip's colours
    std::vector<int32_t> strip_nptr;       // (n_parts * n_strips + 1) ranges into strip_nbr
    std::vector<int32_t> strip_nbr;        // strips (global strip ids, other parts included) a strip shares an edge with
    std::vector<uint8_t> strip_peers;      // (n_parts * n_strips) bit q: part q reads rows of this strip
    int max_strip_nbr = 0;
    std::vector<int32_t> iedge_ptr, gedge_ptr, bcell_ptr;   // (n_parts+1) owned internal edges / ghost edges (index into the
                                      // ghost block) / boundary cells: an internal edge belongs to the part of its lower cell
    int max_row_len = 0;
    int64_t bandwidth = 0;            // max |row - col| after reordering (diagnostic)
};

// Returns an empty string on success, else an error message.
// n_colors > 0: rows are regrouped into that many colours for the multicolour Gauss-Seidel sweeps; hint
// (may be NULL): one signed flow per edge (reference order, > 0 = out of f1) the colours are aligned with.
// n_parts: strips of a domain decomposition (1 = none); rows are then ordered part-major.
// n_strips > 0 (needs n_colors > 0): rows ordered (part, strip, colour, RCM position), see Topology.
// strip_cap > 0: rows a (strip, colour) holds beyond strip_cap move to another legal colour with room (best effort).
std::string build_topology(int n_real, int n_face, int n_edge, const int32_t* f1, const int32_t* f2,
                           bool rcm, int n_colors, const float* hint, int n_parts, Topology& out,
                           int n_strips = 0, int strip_cap = 0);

}  // namespace cwr
