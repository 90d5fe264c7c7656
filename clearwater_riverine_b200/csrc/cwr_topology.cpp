#include "cwr_topology.h"
#include <chrono>
#include <cstdio>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

namespace cwr {
namespace {

// Breadth-first level structure from `start` inside the unvisited component; returns the last
// level's minimum-degree node and the number of levels.  `mark` is scratch (stamp-based).
struct Bfs {
    const std::vector<int32_t>& ptr;
    const std::vector<int32_t>& adj;
    std::vector<int32_t> stamp;
    std::vector<int32_t> queue;
    int32_t cur = 0;
    Bfs(const std::vector<int32_t>& p, const std::vector<int32_t>& a, int n) : ptr(p), adj(a), stamp(n, 0) { queue.reserve(n); }

    int levels(int32_t start, int32_t& far_node) {
        ++cur;
        queue.clear();
        queue.push_back(start);
        stamp[start] = cur;
        size_t head = 0;
        int nlev = 0;
        size_t level_begin = 0;
        while (head < queue.size()) {
            size_t level_end = queue.size();
            level_begin = head;
            for (; head < level_end; ++head) {
                // (the walk is a chain of cache misses -- row pointer, neighbour list, stamps: the queue says which cells
                // come next, so their lines are requested a few cells ahead)
                if (head + 16 < queue.size()) __builtin_prefetch(&ptr[queue[head + 16]]);
                if (head + 8 < queue.size()) __builtin_prefetch(&adj[ptr[queue[head + 8]]]);
                if (head + 4 < queue.size()) {
                    const int32_t w = queue[head + 4];
                    for (int32_t j = ptr[w]; j < ptr[w + 1]; ++j) __builtin_prefetch(&stamp[adj[j]]);
                }
                int32_t u = queue[head];
                for (int32_t j = ptr[u]; j < ptr[u + 1]; ++j) {
                    int32_t v = adj[j];
                    if (stamp[v] != cur) { stamp[v] = cur; queue.push_back(v); }
                }
            }
            ++nlev;
        }
        far_node = queue[level_begin];
        int32_t best = ptr[far_node + 1] - ptr[far_node];
        for (size_t i = level_begin; i < queue.size(); ++i) {
            int32_t d = ptr[queue[i] + 1] - ptr[queue[i]];
            if (d < best) { best = d; far_node = queue[i]; }
        }
        return nlev;
    }
};

// Host threads for the loops of the set-up whose result does not depend on the order of their iterations (counts and
// maxima through atomics, element-wise maps): CWR_TOPO_THREADS, default min(4, hardware threads) -- several ranks of one
// node build their topologies at the same time.  The order-dependent passes (breadth-first walks, stable scatters) stay
// on one thread, so the topology is the same bit for bit with any thread count.
int topo_threads(int64_t work) {
    if (work < (1 << 18)) return 1;
    if (const char* ev = std::getenv("CWR_TOPO_THREADS")) return std::max(1, std::min(64, std::atoi(ev)));
    return (int)std::max(1u, std::min(4u, std::thread::hardware_concurrency()));
}
template <typename F>
void parallel_chunks(int64_t count, F body) {          // body(begin, end) over [0, count) cut into one chunk per thread
    const int T = topo_threads(count);
    if (T <= 1) { body((int64_t)0, count); return; }
    std::vector<std::thread> pool;
    const int64_t per = (count + T - 1) / T;
    bool failed = false;
    int64_t done = 0;
    for (int t = 0; t < T - 1 && !failed; ++t) {
        const int64_t lo = t * per, hi = std::min(count, lo + per);
        try { pool.emplace_back([=, &body] { body(lo, hi); }); done = hi; } catch (...) { failed = true; }
    }
    body(done, count);                                  // (the caller's thread takes the rest: all of it if no thread started)
    for (auto& th : pool) th.join();
}
inline void atomic_inc(int32_t& x) { __atomic_fetch_add(&x, 1, __ATOMIC_RELAXED); }
inline void atomic_max_nonneg(float& x, float q) {      // q >= 0: the bit patterns of non-negative floats order like integers
    int32_t* px = reinterpret_cast<int32_t*>(&x);
    int32_t qi; std::memcpy(&qi, &q, 4);
    int32_t cur = __atomic_load_n(px, __ATOMIC_RELAXED);
    while (cur < qi && !__atomic_compare_exchange_n(px, &cur, qi, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
}

// fraction of a cell's strongest face flow below which an edge does not direct the colouring (0 = every edge
// does); CWR_HINT_TAU overrides it for experiments
float hint_threshold() {
    if (const char* ev = std::getenv("CWR_HINT_TAU")) return (float)std::atof(ev);
    return 0.1f;      // 1M x 16 benchmark, 5 sweeps per application: 3.81 -> 3.14 ms/step (0.25: 3.20)
}

}  // namespace

std::string build_topology(int n_real, int n_face, int n_edge, const int32_t* f1, const int32_t* f2,
                           bool rcm, int n_colors, const float* hint, int n_parts, Topology& T,
                           int n_strips, int strip_cap) {
    if (n_parts < 1 || n_parts > 8) return "n_parts must be in [1, 8]";
    if (n_strips < 0 || n_strips > 4096) return "n_strips must be in [0, 4096]";
    if (n_strips > 0 && n_colors <= 0) return "strips need colours";
    if (n_real <= 0 || n_face < n_real || n_edge <= 0) return "n_real, n_face, n_edge must be positive and n_face >= n_real";
    const int n = n_real, F = n_face, E = n_edge;
    // CWR_TOPO_TIMING=1: seconds per section on stderr (set-up is host work: 16M cells take tens of seconds)
    const bool timing = std::getenv("CWR_TOPO_TIMING") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto tick = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[cwr topology] %-28s %.3f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    int32_t max_f1 = -1;
    int E_int = 0;
    for (int e = 0; e < E; ++e) {
        if (f1[e] < 0 || f1[e] >= n) return "edges_face1 must be a real cell (the reference defines nreal = max(edges_face1))";
        if (f2[e] < 0 || f2[e] >= F) return "edges_face2 out of range";
        if (f1[e] == f2[e]) return "edge connects a cell to itself";
        max_f1 = std::max(max_f1, f1[e]);
        if (f2[e] < n) ++E_int;
    }
    if (max_f1 != n - 1) return "n_real must equal max(edges_face1) + 1 (reference io/hdf.py:268-269)";
    T.n = n; T.F = F; T.E = E; T.E_int = E_int; T.E_g = E - E_int; T.G = F - n;
    T.nnz = 2 * (int64_t)E_int;
    if (T.nnz > 0x7fffffffLL) return "too many non-zeros for 32-bit indices";
    if ((int64_t)E > 0x3fffffffLL) return "too many edges";
    if ((int64_t)n > 0x3fffffffLL) return "too many cells (row indices carry two flag bits)";

    // ---- adjacency of real cells (both directions of every internal edge) --------------------
    std::vector<int32_t> aptr(n + 1, 0), adj(T.nnz);
    // (edge loops touch their two cells' entries at random: the lines are requested kPF edges ahead)
    constexpr int kPF = 16;
    parallel_chunks(E, [&](int64_t lo, int64_t hi) {
        for (int64_t e = lo; e < hi; ++e) {
            if (e + kPF < hi) { __builtin_prefetch(&aptr[f1[e + kPF] + 1], 1); if (f2[e + kPF] < n) __builtin_prefetch(&aptr[f2[e + kPF] + 1], 1); }
            if (f2[e] < n) { atomic_inc(aptr[f1[e] + 1]); atomic_inc(aptr[f2[e] + 1]); }
        }
    });
    for (int i = 0; i < n; ++i) aptr[i + 1] += aptr[i];
    {
        std::vector<int32_t> fill(aptr.begin(), aptr.end() - 1);
        for (int e = 0; e < E; ++e) {
            if (e + kPF < E) { __builtin_prefetch(&fill[f1[e + kPF]], 1); if (f2[e + kPF] < n) __builtin_prefetch(&fill[f2[e + kPF]], 1); }
            if (e + kPF / 2 < E && f2[e + kPF / 2] < n) { __builtin_prefetch(&adj[fill[f1[e + kPF / 2]]], 1); __builtin_prefetch(&adj[fill[f2[e + kPF / 2]]], 1); }
            if (f2[e] < n) { adj[fill[f1[e]]++] = f2[e]; adj[fill[f2[e]]++] = f1[e]; }
        }
    }

    tick("adjacency");
    // ---- reverse Cuthill-McKee -------------------------------------------------------------------
    T.old_of_new.resize(n);
    T.new_of_old.resize(n);
    if (!rcm) {
        std::iota(T.old_of_new.begin(), T.old_of_new.end(), 0);
    } else {
        // cells by (degree, id): a stable counting sort (a comparison sort of 16M cells was ~2 s of the set-up)
        std::vector<int32_t> by_degree(n);
        auto deg = [&](int32_t u) { return aptr[u + 1] - aptr[u]; };
        {
            int max_deg = 0;
            for (int i = 0; i < n; ++i) max_deg = std::max(max_deg, deg(i));
            std::vector<int32_t> dptr((size_t)max_deg + 2, 0);
            for (int i = 0; i < n; ++i) ++dptr[deg(i) + 1];
            for (int d = 0; d <= max_deg; ++d) dptr[d + 1] += dptr[d];
            for (int i = 0; i < n; ++i) by_degree[dptr[deg(i)]++] = i;
        }
        std::vector<uint8_t> visited(n, 0);
        std::vector<int32_t> order;
        order.reserve(n);
        Bfs bfs(aptr, adj, n);
        std::vector<int32_t> nbrs;
        for (int32_t s0 : by_degree) {
            if (visited[s0]) continue;
            // pseudo-peripheral start (George-Liu): walk to the far end while the level count grows
            int32_t start = s0, far = s0;
            int nlev = bfs.levels(start, far);
            for (int it = 0; it < 4; ++it) {
                int32_t far2;
                int nlev2 = bfs.levels(far, far2);
                if (nlev2 <= nlev) break;
                start = far; far = far2; nlev = nlev2;
            }
            size_t head = order.size();
            order.push_back(start);
            visited[start] = 1;
            while (head < order.size()) {
                if (head + 16 < order.size()) __builtin_prefetch(&aptr[order[head + 16]]);
                if (head + 8 < order.size()) __builtin_prefetch(&adj[aptr[order[head + 8]]]);
                if (head + 4 < order.size()) {
                    const int32_t w = order[head + 4];
                    for (int32_t j = aptr[w]; j < aptr[w + 1]; ++j) { __builtin_prefetch(&visited[adj[j]]); __builtin_prefetch(&aptr[adj[j]]); }
                }
                int32_t u = order[head++];
                nbrs.clear();
                for (int32_t j = aptr[u]; j < aptr[u + 1]; ++j) {
                    int32_t v = adj[j];
                    if (!visited[v]) { visited[v] = 1; nbrs.push_back(v); }
                }
                std::sort(nbrs.begin(), nbrs.end(), [&](int32_t a, int32_t b) {
                    int da = deg(a), db = deg(b);
                    return da != db ? da < db : a < b;
                });
                order.insert(order.end(), nbrs.begin(), nbrs.end());
            }
        }
        for (int i = 0; i < n; ++i) T.old_of_new[i] = order[n - 1 - i];
    }
    tick("reverse Cuthill-McKee");
    // ---- flow-aligned multicolouring (for the Gauss-Seidel preconditioner) -------------------------------
    // A sweep visits the colours in order and updates the rows of one colour in parallel, so colours
    // must separate coupled rows.  Upwind advection makes the matrix nearly triangular in the downstream
    // order: a row mostly depends on its UPSTREAM neighbours.  If colours increase along the flow, one
    // sweep carries information across as many cells as there are colours (a row sees the values its
    // upstream neighbours got earlier in the same sweep); colours that ignore the flow carry it ~2 cells.
    // So: level(v) = longest downstream path to v in the flow graph of the hint (Kahn's algorithm; cycles
    // are cut at the node that comes first in RCM order), bumped until (level mod n_colors) differs from
    // every neighbour already coloured; colour = level mod n_colors.  Without a hint every edge is
    // undirected and this is the greedy colouring in RCM order.  Rows are regrouped colour-major, inside
    // a colour by (level, RCM position): one colour = one contiguous row range.
    T.color_ptr.assign(1, 0);
    if (n_colors > 0) {
        int max_deg = 0;
        for (int i = 0; i < n; ++i) max_deg = std::max(max_deg, aptr[i + 1] - aptr[i]);
        if (max_deg + 1 > 64) return "a cell has more than 63 neighbours: multicolour sweeps not available (precond_sweep = 0)";
        const int nc = std::min(64, std::max(n_colors, max_deg + 1));
        std::vector<int32_t> rcm_pos(n);
        for (int i = 0; i < n; ++i) rcm_pos[T.old_of_new[i]] = i;
        // directed flow graph among real cells: up -> down
        // Edges whose flow is weak next to the strongest flow through either of their cells (cross-stream
        // exchange) are left undirected: they carry little coupling, but as graph edges they chain the
        // levels across streamlines (several levels per cell instead of one), so the colours would wrap
        // after two or three cells.  Measured (100k cells, 11 colours): error factor per sweep 0.31 -> 0.25.
        std::vector<int32_t> indeg(n, 0), optr(n + 1, 0), oadj;
        if (hint) {
            const float tau = hint_threshold();
            std::vector<float> cellmax(n, 0.f);
            if (tau > 0.f)
                parallel_chunks(E, [&](int64_t lo, int64_t hi) {
                    for (int64_t e = lo; e < hi; ++e) {
                        if (e + kPF < hi) { __builtin_prefetch(&cellmax[f1[e + kPF]], 1); if (f2[e + kPF] < n) __builtin_prefetch(&cellmax[f2[e + kPF]], 1); }
                        const float q = std::fabs(hint[e]);
                        if (!(q == q)) continue;
                        atomic_max_nonneg(cellmax[f1[e]], q);
                        if (f2[e] < n) atomic_max_nonneg(cellmax[f2[e]], q);
                    }
                });
            auto directed = [&](int e) {
                if (f2[e] >= n || !(hint[e] != 0.f) || hint[e] != hint[e]) return false;
                return tau <= 0.f || std::fabs(hint[e]) >= tau * std::max(cellmax[f1[e]], cellmax[f2[e]]);
            };
            // (classified once: 0 = undirected, 1 = f1 -> f2, 2 = f2 -> f1)
            std::vector<uint8_t> dir(E);
            parallel_chunks(E, [&](int64_t lo, int64_t hi) {
                for (int64_t e = lo; e < hi; ++e) {
                    if (e + kPF < hi && f2[e + kPF] < n) { __builtin_prefetch(&cellmax[f1[e + kPF]]); __builtin_prefetch(&cellmax[f2[e + kPF]]); }
                    dir[e] = !directed((int)e) ? 0 : (hint[e] > 0.f ? 1 : 2);
                }
                for (int64_t e = lo; e < hi; ++e) {
                    if (e + kPF < hi && dir[e + kPF]) { __builtin_prefetch(&optr[f1[e + kPF] + 1], 1); __builtin_prefetch(&optr[f2[e + kPF] + 1], 1);
                                                        __builtin_prefetch(&indeg[f1[e + kPF]], 1); __builtin_prefetch(&indeg[f2[e + kPF]], 1); }
                    if (!dir[e]) continue;
                    const int32_t up = dir[e] == 1 ? f1[e] : f2[e], down = dir[e] == 1 ? f2[e] : f1[e];
                    atomic_inc(optr[up + 1]); atomic_inc(indeg[down]);
                }
            });
            for (int i = 0; i < n; ++i) optr[i + 1] += optr[i];
            oadj.resize(optr[n]);
            std::vector<int32_t> fill(optr.begin(), optr.end() - 1);
            for (int e = 0; e < E; ++e) {
                if (e + kPF < E && dir[e + kPF]) { __builtin_prefetch(&fill[f1[e + kPF]], 1); __builtin_prefetch(&fill[f2[e + kPF]], 1); }
                if (!dir[e]) continue;
                const int32_t up = dir[e] == 1 ? f1[e] : f2[e], down = dir[e] == 1 ? f2[e] : f1[e];
                oadj[fill[up]++] = down;
            }
        }
        std::vector<int32_t> level(n, -1), tentative(n, 0), queue;
        queue.reserve(n);
        for (int i = 0; i < n; ++i) { const int32_t u = T.old_of_new[i]; if (indeg[u] == 0) queue.push_back(u); }
        std::vector<uint8_t> queued(n, 0);
        for (int32_t u : queue) queued[u] = 1;
        size_t head = 0;
        int scan = 0;                 // next RCM position to look at when a cycle blocks the queue
        int max_level = 0;
        for (int done = 0; done < n; ++done) {
            if (head == queue.size()) {            // only cycles left: cut one at the first unprocessed node
                while (queued[T.old_of_new[scan]]) ++scan;
                const int32_t u = T.old_of_new[scan];
                queued[u] = 1; queue.push_back(u);
            }
            if (head + 16 < queue.size()) { __builtin_prefetch(&aptr[queue[head + 16]]); __builtin_prefetch(&optr[queue[head + 16]]); }
            if (head + 8 < queue.size()) {
                const int32_t w = queue[head + 8];
                __builtin_prefetch(&adj[aptr[w]]);
                if (!oadj.empty()) __builtin_prefetch(&oadj[std::min<size_t>(optr[w], oadj.size() - 1)]);
                __builtin_prefetch(&tentative[w]);
            }
            if (head + 4 < queue.size()) {
                const int32_t w = queue[head + 4];
                for (int32_t j = aptr[w]; j < aptr[w + 1]; ++j) __builtin_prefetch(&level[adj[j]]);
                for (int32_t j = optr[w]; j < optr[w + 1]; ++j) { __builtin_prefetch(&tentative[oadj[j]]); __builtin_prefetch(&indeg[oadj[j]]); }
            }
            const int32_t u = queue[head++];
            uint64_t used = 0;
            for (int32_t j = aptr[u]; j < aptr[u + 1]; ++j) {
                const int32_t lv = level[adj[j]];
                if (lv >= 0) used |= (uint64_t)1 << (lv % nc);
            }
            int lv = tentative[u];
            for (int tries = 0; tries < nc && ((used >> (lv % nc)) & 1); ++tries) ++lv;   // degree < nc: always found
            level[u] = lv;
            max_level = std::max(max_level, lv);
            for (int32_t j = optr[u]; j < optr[u + 1]; ++j) {
                const int32_t v = oadj[j];
                tentative[v] = std::max(tentative[v], lv + 1);
                if (--indeg[v] == 0 && !queued[v]) { queued[v] = 1; queue.push_back(v); }
            }
        }
        tick("levels / colours");
        // ---- partition into n_parts strips (domain decomposition), then colour-major inside a part ------
        // Parts are equal chunks of the RCM order (strips of the RCM band: breadth-first wavefronts, so a
        // part only couples to its two neighbouring strips and the cut is a smooth front).  Final order:
        // (part, colour, level, RCM position), sorted through one 64-bit key per cell.
        if (max_level >= (1 << 22)) return "flow hint has more than 4M downstream levels";
        std::vector<int32_t> part(n);
        for (int i = 0; i < n; ++i) part[i] = (int32_t)(((int64_t)rcm_pos[i] * n_parts) / n);
        // strips: equal chunks of the RCM order nested in the parts (global strip id = part * n_strips + strip)
        std::vector<int32_t> strip;
        if (n_strips > 0) {
            strip.resize(n);
            for (int i = 0; i < n; ++i) strip[i] = (int32_t)(((int64_t)rcm_pos[i] * n_parts * n_strips) / n);
        }
        std::vector<uint8_t> colr(n);
        for (int i = 0; i < n; ++i) colr[i] = (uint8_t)(level[i] % nc);
        // ---- strips: no (strip, colour) above strip_cap rows -------------------------------------------------
        // The sweep kernel handles strip_cap rows of a colour per pass; the level colours are uneven (1M-cell
        // benchmark, 16 colours: mean 211 rows, max 307) and the strips march in lock-step with their neighbours,
        // so one over-full colour slows every strip around it.  The rows an over-full colour holds beyond the cap
        // (last in RCM order) move to the nearest later colour no neighbour has and that still has room: they are
        // swept a little later in the sweep than their level asks for -- still after their upstream neighbours.
        // (without strips, on one part: the whole mesh is the one strip -- the on-chip solver k_solve_chip takes a colour
        // in one pass of its CTA)
        const bool one_strip = n_strips == 0 && n_parts == 1 && strip_cap > 0;
        if (one_strip) strip.assign(n, 0);
        if ((n_strips > 0 || one_strip) && strip_cap > 0) {
            const int NS = one_strip ? 1 : n_parts * n_strips;
            std::vector<int32_t> cnt((size_t)NS * nc, 0);
            for (int i = 0; i < n; ++i) ++cnt[(size_t)strip[i] * nc + colr[i]];
            // (a strip whose rows do not fit n_colors passes with 8 % to spare is balanced towards 1.25 x its mean colour
            // instead: moving rows just to fail the cap anyway only costs sweeps -- 16M cells x 1: 18 per step against 13)
            std::vector<int32_t> cap_of(NS, strip_cap);
            {
                std::vector<int64_t> rows(NS, 0);
                for (int i = 0; i < n; ++i) ++rows[strip[i]];
                for (int sidx = 0; sidx < NS; ++sidx)
                    if ((int64_t)nc * strip_cap * 100 < rows[sidx] * (one_strip ? 103 : 108))      // (on chip 3 % to spare will do)
                        cap_of[sidx] = std::max<int64_t>(strip_cap, (rows[sidx] * 5 + 4 * nc - 1) / (4 * nc));
            }
            for (int pos = n - 1; pos >= 0; --pos) {
                const int32_t u = T.old_of_new[pos];
                int32_t* cs = cnt.data() + (size_t)strip[u] * nc;
                const int32_t strip_cap = cap_of[strip[u]];
                if (cs[colr[u]] <= strip_cap) continue;
                uint64_t used = 0;
                for (int32_t j = aptr[u]; j < aptr[u + 1]; ++j) used |= (uint64_t)1 << colr[adj[j]];
                for (int d = 1; d < nc; ++d) {
                    const int c2 = (colr[u] + d) % nc;
                    if (((used >> c2) & 1) || cs[c2] >= strip_cap) continue;
                    --cs[colr[u]]; ++cs[c2];
                    colr[u] = (uint8_t)c2;
                    break;
                }
            }
        }
        if (one_strip) std::vector<int32_t>().swap(strip);
        // (stable counting sorts over the RCM order instead of one comparison sort of n 64-bit keys: the row order is
        // most of the set-up's sorting at 16M cells)
        std::vector<int32_t> order(n);
        {
            const bool by_strip = n_strips > 0;
            std::vector<int32_t> seq(n);               // cells in RCM order, then (non-strip mode) stably by level
            if (by_strip) seq = T.old_of_new;
            else {
                std::vector<int32_t> lptr((size_t)max_level + 2, 0);
                for (int i = 0; i < n; ++i) ++lptr[level[i] + 1];
                for (int l = 0; l <= max_level; ++l) lptr[l + 1] += lptr[l];
                for (int pos = 0; pos < n; ++pos) { const int32_t u = T.old_of_new[pos]; seq[lptr[level[u]]++] = u; }
            }
            const size_t nb = (size_t)(by_strip ? n_parts * n_strips : n_parts) * nc;
            std::vector<int32_t> bptr(nb + 1, 0);
            auto bucket = [&](int32_t u) { return (size_t)(by_strip ? strip[u] : part[u]) * nc + colr[u]; };
            for (int i = 0; i < n; ++i) ++bptr[bucket(i) + 1];
            for (size_t b = 0; b < nb; ++b) bptr[b + 1] += bptr[b];
            for (int q = 0; q < n; ++q) { const int32_t u = seq[q]; order[bptr[bucket(u)]++] = u; }
        }
        T.n_colors = nc;
        T.color_ptr.assign((size_t)n_parts * (nc + 1), 0);
        T.part_ptr.assign(n_parts + 1, 0);
        {
            std::vector<int32_t> cnt((size_t)n_parts * nc, 0);
            for (int i = 0; i < n; ++i) ++cnt[(size_t)part[i] * nc + colr[i]];
            int32_t pos = 0;
            for (int p = 0; p < n_parts; ++p) {
                T.part_ptr[p] = pos;
                for (int c = 0; c < nc; ++c) { T.color_ptr[(size_t)p * (nc + 1) + c] = pos; pos += cnt[(size_t)p * nc + c]; }
                T.color_ptr[(size_t)p * (nc + 1) + nc] = pos;
            }
            T.part_ptr[n_parts] = pos;
        }
        T.n_strips = n_strips;
        T.strip_cptr.clear();
        if (n_strips > 0) {        // rows are (strip, colour)-major: the ranges of every strip's colours
            const int NS = n_parts * n_strips;
            std::vector<int32_t> cnt((size_t)NS * nc, 0);
            for (int i = 0; i < n; ++i) ++cnt[(size_t)strip[i] * nc + colr[i]];
            T.strip_cptr.assign((size_t)NS * (nc + 1), 0);
            int32_t pos = 0;
            for (int s = 0; s < NS; ++s) {
                for (int c = 0; c < nc; ++c) { T.strip_cptr[(size_t)s * (nc + 1) + c] = pos; pos += cnt[(size_t)s * nc + c]; }
                T.strip_cptr[(size_t)s * (nc + 1) + nc] = pos;
            }
            // color_ptr keeps its shape; a part's colours are not contiguous in this mode
            for (int p = 0; p < n_parts; ++p)
                for (int c = 0; c <= nc; ++c) T.color_ptr[(size_t)p * (nc + 1) + c] = c == 0 ? T.part_ptr[p] : T.part_ptr[p + 1];
        }
        T.old_of_new.swap(order);
        T.n_levels = max_level + 1;
        T.color_of.resize(n);
        for (int i = 0; i < n; ++i) T.color_of[i] = colr[T.old_of_new[i]];       // by new id
    } else {
        // no colours: parts are equal chunks of the RCM order
        T.n_colors = 0;
        T.color_ptr.clear();
        T.part_ptr.assign(n_parts + 1, 0);
        for (int p = 0; p <= n_parts; ++p) T.part_ptr[p] = (int32_t)(((int64_t)p * n) / n_parts);
        T.color_of.clear();
    }
    for (int i = 0; i < n; ++i) T.new_of_old[T.old_of_new[i]] = i;

    tick("strip balancing / row order");
    // ---- edge renumbering ---------------------------------------------------------------------------
    // internal edges: sort by (min new cell, max new cell, original id); ghost edges: by (new cell, original id)
    std::vector<int32_t> internal, ghost, ea(E), eb(E);       // ea / eb: new ids of an edge's cells (eb = -1: ghost cell)
    {
        // counting sort by the lower new cell (edges scattered in original order: ties keep ascending ids), then the
        // handful of edges of a cell by (higher cell, id)
        // (the new ids of an edge's cells are random reads of new_of_old: looked up once, kept for the second pass)
        std::vector<int32_t> iptr(n + 1, 0), gptr(n + 1, 0);
        parallel_chunks(E, [&](int64_t lo, int64_t hi) {
            for (int64_t e = lo; e < hi; ++e) {
                if (e + kPF < hi) { __builtin_prefetch(&T.new_of_old[f1[e + kPF]]); if (f2[e + kPF] < n) __builtin_prefetch(&T.new_of_old[f2[e + kPF]]); }
                const int32_t a = T.new_of_old[f1[e]];
                ea[e] = a;
                if (f2[e] < n) { eb[e] = T.new_of_old[f2[e]]; atomic_inc(iptr[std::min(a, eb[e]) + 1]); }
                else { eb[e] = -1; atomic_inc(gptr[a + 1]); }
            }
        });
        for (int i = 0; i < n; ++i) { iptr[i + 1] += iptr[i]; gptr[i + 1] += gptr[i]; }
        internal.resize(E_int); ghost.resize(T.E_g);
        std::vector<int32_t> hi(E_int);
        {
            std::vector<int32_t> ifill(iptr.begin(), iptr.end() - 1), gfill(gptr.begin(), gptr.end() - 1);
            for (int e = 0; e < E; ++e) {
                if (e + kPF < E) __builtin_prefetch(&ifill[eb[e + kPF] >= 0 ? std::min(ea[e + kPF], eb[e + kPF]) : 0], 1);
                if (e + kPF / 2 < E && eb[e + kPF / 2] >= 0) {
                    const int32_t o = ifill[std::min(ea[e + kPF / 2], eb[e + kPF / 2])];
                    __builtin_prefetch(&internal[std::min(o, E_int - 1)], 1); __builtin_prefetch(&hi[std::min(o, E_int - 1)], 1);
                }
                const int32_t a = ea[e], b = eb[e];
                if (b >= 0) {
                    const int32_t o = ifill[std::min(a, b)]++;
                    internal[o] = e; hi[o] = std::max(a, b);
                } else ghost[gfill[a]++] = e;
            }
        }
        for (int i = 0; i < n; ++i)
            for (int32_t j = iptr[i] + 1; j < iptr[i + 1]; ++j) {          // insertion sort, stable
                const int32_t e = internal[j], h = hi[j];
                int32_t q = j;
                while (q > iptr[i] && hi[q - 1] > h) { internal[q] = internal[q - 1]; hi[q] = hi[q - 1]; --q; }
                internal[q] = e; hi[q] = h;
            }
    }
    T.eperm.resize(E); T.f1p.resize(E); T.f2p.resize(E);
    for (int i = 0; i < E_int; ++i) T.eperm[i] = internal[i];
    for (int i = 0; i < T.E_g; ++i) T.eperm[E_int + i] = ghost[i];
    parallel_chunks(E, [&](int64_t lo, int64_t hi) {
        for (int64_t ep = lo; ep < hi; ++ep) {
            if (ep + kPF < hi) { __builtin_prefetch(&ea[T.eperm[ep + kPF]]); __builtin_prefetch(&eb[T.eperm[ep + kPF]]); }
            const int32_t e = T.eperm[ep];
            T.f1p[ep] = ea[e];
            T.f2p[ep] = eb[e] >= 0 ? eb[e] : f2[e];            // ghost cells keep their id (>= n)
        }
    });

    tick("edge renumbering");
    // ---- off-diagonal CSR with slot -> (edge, side) --------------------------------------------------
    T.rowptr.assign(n + 1, 0);
    for (int ep = 0; ep < E_int; ++ep) { ++T.rowptr[T.f1p[ep] + 1]; ++T.rowptr[T.f2p[ep] + 1]; }
    for (int i = 0; i < n; ++i) T.rowptr[i + 1] += T.rowptr[i];
    T.col.resize(T.nnz); T.slot_edge.resize(T.nnz);
    {
        // (col, slot code) pairs per row, then sort each row by column
        std::vector<int32_t> fill(T.rowptr.begin(), T.rowptr.end() - 1);
        for (int ep = 0; ep < E_int; ++ep) {
            int32_t P = T.f1p[ep], N = T.f2p[ep];
            int32_t a = fill[P]++; T.col[a] = N; T.slot_edge[a] = (ep << 1) | 0;   // A[P,N]
            int32_t b = fill[N]++; T.col[b] = P; T.slot_edge[b] = (ep << 1) | 1;   // A[N,P]
        }
        T.max_row_len = 0; T.bandwidth = 0;
        const int NT = topo_threads(n);
        std::vector<int32_t> len_of(NT + 1, 0);
        std::vector<int64_t> band_of(NT + 1, 0);
        const int64_t per = ((int64_t)n + NT - 1) / NT;
        parallel_chunks(n, [&](int64_t lo, int64_t hi) {        // rows are independent; maxima per chunk, merged below
            std::vector<std::pair<int32_t, int32_t>> tmp;
            int32_t len = 0; int64_t band = 0;
            for (int64_t i = lo; i < hi; ++i) {
                int32_t s = T.rowptr[i], e = T.rowptr[i + 1];
                len = std::max(len, e - s);
                tmp.clear();
                for (int32_t j = s; j < e; ++j) tmp.emplace_back(T.col[j], T.slot_edge[j]);
                std::sort(tmp.begin(), tmp.end());
                for (int32_t j = s; j < e; ++j) {
                    T.col[j] = tmp[j - s].first; T.slot_edge[j] = tmp[j - s].second;
                    band = std::max<int64_t>(band, std::abs((int64_t)T.col[j] - i));
                }
            }
            const int slot = (int)std::min<int64_t>(NT, lo / std::max<int64_t>(1, per));   // (chunks start at multiples of per, or at 0)
            len_of[slot] = std::max(len_of[slot], len); band_of[slot] = std::max(band_of[slot], band);
        });
        for (int t = 0; t <= NT; ++t) { T.max_row_len = std::max(T.max_row_len, len_of[t]); T.bandwidth = std::max(T.bandwidth, band_of[t]); }
    }

    tick("CSR pattern");
    // ---- row-major ELL copy of the pattern (what the kernels read) ------------------------------------
    T.W = std::max(4, (T.max_row_len + 3) / 4 * 4);
    T.ell_col.assign((size_t)n * T.W, 0);
    T.ell_code.assign((size_t)n * T.W, -1);
    parallel_chunks(n, [&](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            const int32_t s = T.rowptr[i], e = T.rowptr[i + 1];
            for (int w = 0; w < T.W; ++w) {
                const size_t o = (size_t)i * T.W + w;
                if (s + w < e) { T.ell_col[o] = T.col[s + w]; T.ell_code[o] = T.slot_edge[s + w]; }
                else T.ell_col[o] = (int32_t)i;
            }
        }
    });

    tick("ELL pattern");
    // ---- boundary cells ----------------------------------------------------------------------------------
    T.bcell.clear(); T.bptr.clear(); T.bedge.resize(T.E_g);
    for (int i = 0; i < T.E_g; ++i) {
        int ep = E_int + i;
        if (i == 0 || T.f1p[ep] != T.f1p[ep - 1]) { T.bcell.push_back(T.f1p[ep]); T.bptr.push_back(i); }
        T.bedge[i] = ep;
    }
    T.bptr.push_back(T.E_g);

    // ---- Gauss-Seidel: mark the ELL entries whose neighbour is visited LATER in a sweep ------------------
    // (colour >= the row's colour; padding points at the row itself).  During the first sweep from z = 0
    // those neighbours still hold 0.  Bit 31 of ell_col; every kernel masks it off.
    if (T.n_colors > 0)
        parallel_chunks(n, [&](int64_t lo, int64_t hi) {
            for (int64_t i = lo; i < hi; ++i)
                for (int w = 0; w < T.W; ++w) {
                    int32_t& cj = T.ell_col[(size_t)i * T.W + w];
                    const int ci = T.color_of[i], cn = T.color_of[cj];
                    if (cn >= ci) cj |= kLaterBit;
                    if (cn == (ci + T.n_colors - 1) % T.n_colors && cn != ci) cj |= kPrevBit;
                }
        });

    tick("boundary cells / sweep flags");
    // ---- strips: which strips a strip's rows are coupled to (global strip ids; strips of other parts included:
    // the sweep kernel synchronises with them through flag mirrors in peer memory) ---------------------------
    T.strip_nptr.clear(); T.strip_nbr.clear(); T.max_strip_nbr = 0;
    if (T.n_strips > 0) {
        const int NS = n_parts * T.n_strips, nc = T.n_colors;
        std::vector<int32_t> srow(n);
        for (int s = 0; s < NS; ++s)
            for (int32_t i = T.strip_cptr[(size_t)s * (nc + 1)]; i < T.strip_cptr[(size_t)s * (nc + 1) + nc]; ++i) srow[i] = s;
        std::vector<uint64_t> pairs;
        for (int ep = 0; ep < E_int; ++ep) {
            const int32_t sa = srow[T.f1p[ep]], sb = srow[T.f2p[ep]];
            if (sa == sb) continue;
            pairs.push_back(((uint64_t)(uint32_t)sa << 32) | (uint32_t)sb);
            pairs.push_back(((uint64_t)(uint32_t)sb << 32) | (uint32_t)sa);
        }
        std::sort(pairs.begin(), pairs.end());
        pairs.erase(std::unique(pairs.begin(), pairs.end()), pairs.end());
        T.strip_nptr.assign(NS + 1, 0);
        for (uint64_t pr : pairs) ++T.strip_nptr[(pr >> 32) + 1];
        for (int s = 0; s < NS; ++s) {
            T.max_strip_nbr = std::max(T.max_strip_nbr, T.strip_nptr[s + 1]);
            T.strip_nptr[s + 1] += T.strip_nptr[s];
        }
        T.strip_nbr.reserve(pairs.size());
        for (uint64_t pr : pairs) T.strip_nbr.push_back((int32_t)(uint32_t)pr);
    }

    tick("strip neighbours");
    // ---- domain decomposition: who reads whose rows, which edges / boundary cells a part owns ----------
    const int P = n_parts;
    auto part_of = [&](int32_t row) { return (int)(std::upper_bound(T.part_ptr.begin(), T.part_ptr.end(), row) - T.part_ptr.begin()) - 1; };
    T.send_mask.assign(n, 0);
    if (P > 1)
        for (int ep = 0; ep < E_int; ++ep) {
            const int pa = part_of(T.f1p[ep]), pb = part_of(T.f2p[ep]);
            if (pa != pb) { T.send_mask[T.f1p[ep]] |= (uint8_t)(1u << pb); T.send_mask[T.f2p[ep]] |= (uint8_t)(1u << pa); }
        }
    // the parts that read rows of a strip (its flag is mirrored there)
    T.strip_peers.clear();
    if (T.n_strips > 0) {
        const int NS = n_parts * T.n_strips, nc = T.n_colors;
        T.strip_peers.assign(NS, 0);
        for (int sidx = 0; sidx < NS; ++sidx)
            for (int32_t i = T.strip_cptr[(size_t)sidx * (nc + 1)]; i < T.strip_cptr[(size_t)sidx * (nc + 1) + nc]; ++i)
                T.strip_peers[sidx] |= T.send_mask[i];
    }
    T.send_ptr.assign(P + 1, 0); T.send_rows.clear();
    for (int pp = 0; pp < P; ++pp) {
        T.send_ptr[pp] = (int32_t)T.send_rows.size();
        for (int32_t i = T.part_ptr[pp]; i < T.part_ptr[pp + 1]; ++i) if (T.send_mask[i]) T.send_rows.push_back(i);
    }
    T.send_ptr[P] = (int32_t)T.send_rows.size();
    // internal edges are sorted by their lower cell, ghost edges and boundary cells by their cell: the
    // owner (part of the lower / the real cell) is non-decreasing along each list
    T.iedge_ptr.assign(P + 1, E_int); T.gedge_ptr.assign(P + 1, T.E_g); T.bcell_ptr.assign(P + 1, (int32_t)T.bcell.size());
    {
        int ep = 0, g = 0, b = 0;
        for (int pp = 0; pp < P; ++pp) {
            const int32_t lo_row = T.part_ptr[pp];
            while (ep < E_int && std::min(T.f1p[ep], T.f2p[ep]) < lo_row) ++ep;
            while (g < T.E_g && T.f1p[E_int + g] < lo_row) ++g;
            while (b < (int)T.bcell.size() && T.bcell[b] < lo_row) ++b;
            T.iedge_ptr[pp] = ep; T.gedge_ptr[pp] = g; T.bcell_ptr[pp] = b;
        }
    }
    return "";
}

}  // namespace cwr
