// fem_api.cu -- C ABI of the gel FEM substep (see include/tacex_b200.h).
#include "tx_kernels.h"
#include <new>
#include <string>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <vector>

using namespace tx;

struct tx_fem {
    tx_fem_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    std::vector<double> mass;
    int *d_tets = nullptr, *d_attach = nullptr, *d_surf = nullptr;
    double *d_Dm_inv = nullptr, *d_vol = nullptr, *d_mass = nullptr, *d_X = nullptr, *d_tsc = nullptr, *d_valg = nullptr, *d_xt = nullptr;
    int *d_adj_off = nullptr, *d_adj = nullptr, *d_edge_off = nullptr, *d_edge_adj = nullptr, *d_ell = nullptr;
    int *d_attach_of = nullptr, *d_surf_of = nullptr;
    int nE = 0, n_s = 0, nslots = 0;
    int chunk = 0; // tets per assembly chunk (<= threads): sets the per-CTA scratch footprint in L2
    long long* d_cycles = nullptr;
    int grid = 0;
    // markers
    int M = 0;
    int* d_tri = nullptr;
    double* d_w = nullptr;
    double cam_R[9], cam_t[3], fx = 0, fy = 0, cx = 0, cy = 0;
    int normalize = 0, zero_all = 0;
    double half_w = 160.0;
    // top-surface triangles for the height-map rasteriser
    int n_top = 0;
    int* d_top = nullptr;
    // prescribed triangle-mesh indenter (tx_fem_set_indenter_mesh)
    int mesh_n = 0;
    double *d_mesh_tri = nullptr, *d_mesh_box = nullptr;
    // second half of the vertex-face contact: the mesh's unique vertices against the gel's contact triangles
    int mesh_nv = 0, n_ctri = 0;
    double* d_mesh_vert = nullptr;
    int *d_ctri = nullptr, *d_ctri_row_start = nullptr, *d_ctri_row_adj = nullptr, *d_ctri_edge_start = nullptr, *d_ctri_edge_adj = nullptr;
    std::map<std::pair<int, int>, int> edge_of; // (i, j), i < j -> edge number (tx_fem_create)
    std::vector<double> h_X;                    // rest positions (mollifier threshold of the edge-edge candidates)
    int n_cedge = 0, mesh_ne = 0;
    MeshGrid mgrid[3] = {};     // broad phase over the mesh's triangles / vertices / edges (device pointers inside)
    int* d_grid_buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int *d_cedge = nullptr, *d_cedge_row_start = nullptr, *d_cedge_row_adj = nullptr, *d_edge_cedge = nullptr, *d_mesh_edge = nullptr;
    double* d_cedge_len2 = nullptr;
};

// Default assembly chunk (see fem_kernel.cu, grad_hess): measured on the B200, profiles/r02_fem_chunk.txt
static constexpr int FEM_DEFAULT_CHUNK = 576;
static std::string g_fem_err;
static int ffail(tx_fem* f, int code, const std::string& m)
{
    if (f) f->err = m; else g_fem_err = m;
    return code;
}
#define FEM_CUDA(f, expr)                                                                                             \
    do {                                                                                                              \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess) return ffail((f), TX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

extern "C" const char* tx_fem_last_error(const tx_fem* f) { return f ? f->err.c_str() : g_fem_err.c_str(); }

extern "C" int tx_fem_create(const tx_fem_config* c, const double* X, const int32_t* tets, const int32_t* attach,
                             const int32_t* surf, int device, void* cuda_stream, tx_fem** out)
{
    if (!c || !X || !tets || !out || (c->A > 0 && !attach) || (c->S > 0 && !surf))
        return ffail(nullptr, TX_ERR_INVALID_ARG, "tx_fem_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return ffail(nullptr, TX_ERR_NO_DEVICE, "tx_fem_create: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return ffail(nullptr, TX_ERR_INVALID_ARG, "tx_fem_create: bad device index");
    if (c->V <= 0 || c->T <= 0 || c->substep <= 0) return ffail(nullptr, TX_ERR_INVALID_ARG, "tx_fem_create: bad sizes");
    if (c->V > fem_threads())
        return ffail(nullptr, TX_ERR_UNSUPPORTED, "tx_fem_create: more vertices than the one-row-per-thread solver supports (576)");
    tx_fem* f = new (std::nothrow) tx_fem();
    if (!f) return ffail(nullptr, TX_ERR_INVALID_ARG, "out of host memory");
    f->cfg = *c;
    f->device = device;
    f->stream = (cudaStream_t)cuda_stream;
    f->h_X.assign(X, X + (size_t)3 * c->V);
    {   // tets per assembly chunk: a multiple of 32 in [32, threads]; TX_FEM_CHUNK overrides the default for experiments
        int ch = FEM_DEFAULT_CHUNK;
        if (const char* ev = getenv("TX_FEM_CHUNK")) ch = atoi(ev);
        ch = (ch / 32) * 32;
        f->chunk = ch < 32 ? 32 : (ch > fem_threads() ? fem_threads() : ch);
    }
    // precompute Dm^-1, elastic rest volume, lumped mass (ref: finite_element_method.cu:957-982, 721-744)
    std::vector<double> Dmi((size_t)9 * c->T), vol(c->T);
    f->mass.assign(c->V, 0.0);
    for (int t = 0; t < c->T; ++t) {
        const int32_t* e = tets + 4 * t;
        for (int k = 0; k < 4; ++k)
            if (e[k] < 0 || e[k] >= c->V) { delete f; return ffail(nullptr, TX_ERR_INVALID_ARG, "tx_fem_create: tet index out of range"); }
        double Dm[9];
        for (int a = 0; a < 3; ++a)
            for (int k = 0; k < 3; ++k) Dm[a * 3 + k] = X[3 * e[k + 1] + a] - X[3 * e[0] + a];
        const double det = Dm[0] * (Dm[4] * Dm[8] - Dm[5] * Dm[7]) - Dm[1] * (Dm[3] * Dm[8] - Dm[5] * Dm[6]) +
                           Dm[2] * (Dm[3] * Dm[7] - Dm[4] * Dm[6]);
        if (!(det > 0.0)) { delete f; return ffail(nullptr, TX_ERR_INVALID_ARG, "tx_fem_create: inverted or degenerate tet"); }
        double* B = Dmi.data() + 9 * t;
        B[0] = (Dm[4] * Dm[8] - Dm[5] * Dm[7]) / det; B[1] = (Dm[2] * Dm[7] - Dm[1] * Dm[8]) / det; B[2] = (Dm[1] * Dm[5] - Dm[2] * Dm[4]) / det;
        B[3] = (Dm[5] * Dm[6] - Dm[3] * Dm[8]) / det; B[4] = (Dm[0] * Dm[8] - Dm[2] * Dm[6]) / det; B[5] = (Dm[2] * Dm[3] - Dm[0] * Dm[5]) / det;
        B[6] = (Dm[3] * Dm[7] - Dm[4] * Dm[6]) / det; B[7] = (Dm[1] * Dm[6] - Dm[0] * Dm[7]) / det; B[8] = (Dm[0] * Dm[4] - Dm[1] * Dm[3]) / det;
        vol[t] = c->rest_volume_det ? det : det / 6.0;
        for (int k = 0; k < 4; ++k) f->mass[e[k]] += c->density * (det / 6.0) / 4.0;
    }
#define FEM_CUDA_C(expr)                                                                                              \
    do {                                                                                                              \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess) {                                                                                      \
            ffail(nullptr, TX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                          \
            tx_fem_destroy(f);                                                                                        \
            return TX_ERR_CUDA;                                                                                       \
        }                                                                                                             \
    } while (0)
    FEM_CUDA_C(cudaSetDevice(device));
    cudaDeviceProp prop;
    FEM_CUDA_C(cudaGetDeviceProperties(&prop, device));
    f->grid = prop.multiProcessorCount; // one persistent CTA per SM
    FEM_CUDA_C(cudaMalloc(&f->d_tets, sizeof(int) * 4 * c->T));
    FEM_CUDA_C(cudaMalloc(&f->d_attach, sizeof(int) * (c->A > 0 ? c->A : 1)));
    FEM_CUDA_C(cudaMalloc(&f->d_surf, sizeof(int) * (c->S > 0 ? c->S : 1)));
    FEM_CUDA_C(cudaMalloc(&f->d_Dm_inv, sizeof(double) * 9 * c->T));
    FEM_CUDA_C(cudaMalloc(&f->d_vol, sizeof(double) * c->T));
    FEM_CUDA_C(cudaMalloc(&f->d_mass, sizeof(double) * c->V));
    FEM_CUDA_C(cudaMalloc(&f->d_X, sizeof(double) * 3 * c->V));
    FEM_CUDA_C(cudaMalloc(&f->d_tsc, sizeof(double) * 4 * 30 * (size_t)fem_threads() * f->grid));
    FEM_CUDA_C(cudaMalloc(&f->d_xt, sizeof(double) * 3 * (size_t)c->V * f->grid));
    {   // vertex graph: edges (i < j) numbered by (j - i, i) so that consecutive rows read consecutive blocks; edge ->
        // (tet, pair slot, transpose) incidence in ascending tet order; ELL rows (one slot per distinct offset j - i when
        // the mesh is structured, otherwise the neighbours in ascending order)
        static const int PAIR[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
        std::map<std::pair<int, int>, std::vector<int>> inc; // key (offset, i)
        for (int t = 0; t < c->T; ++t)
            for (int ps = 0; ps < 6; ++ps) {
                const int a = tets[4 * t + PAIR[ps][0]], b = tets[4 * t + PAIR[ps][1]];
                if (a == b) { tx_fem_destroy(f); return ffail(nullptr, TX_ERR_INVALID_ARG, "tx_fem_create: degenerate tet"); }
                const int i = a < b ? a : b, j = a < b ? b : a;
                inc[{j - i, i}].push_back(t << 4 | ps << 1 | (a > b ? 1 : 0));
            }
        const int nE = (int)inc.size();
        std::vector<int> eoff(nE + 1, 0), eadj, ei(nE), ej(nE);
        std::vector<int> offsets; // distinct directed offsets
        {
            int e = 0;
            for (auto& kv : inc) {
                ei[e] = kv.first.second;
                ej[e] = kv.first.second + kv.first.first;
                f->edge_of[{ei[e], ej[e]}] = e;
                for (int en : kv.second) eadj.push_back(en);
                eoff[e + 1] = (int)eadj.size();
                if (offsets.empty() || offsets.back() != kv.first.first) offsets.push_back(kv.first.first);
                ++e;
            }
        }
        const int TH = fem_threads();
        std::vector<std::vector<int>> rows(c->V); // packed entries per row
        int nslots = 0;
        std::vector<int> ell;
        if ((int)offsets.size() <= 16) { // structured: slot = rank of the directed offset (negative offsets first)
            nslots = 2 * (int)offsets.size();
            ell.assign((size_t)nslots * TH, -1);
            for (int e = 0; e < nE; ++e) {
                const int k = (int)(std::lower_bound(offsets.begin(), offsets.end(), ej[e] - ei[e]) - offsets.begin());
                ell[(size_t)(offsets.size() + k) * TH + ei[e]] = ej[e] | (e << 13);            // row i, neighbour j
                ell[(size_t)(offsets.size() - 1 - k) * TH + ej[e]] = ei[e] | 0x1000 | (e << 13); // row j, neighbour i (A^T)
            }
        } else {
            for (int e = 0; e < nE; ++e) {
                rows[ei[e]].push_back(ej[e] | (e << 13));
                rows[ej[e]].push_back(ei[e] | 0x1000 | (e << 13));
            }
            for (auto& r : rows) nslots = std::max(nslots, (int)r.size());
            ell.assign((size_t)nslots * TH, -1);
            for (int i = 0; i < c->V; ++i) {
                std::sort(rows[i].begin(), rows[i].end(), [](int x, int y) { return (x & 0xfff) < (y & 0xfff); });
                for (size_t k = 0; k < rows[i].size(); ++k) ell[k * TH + i] = rows[i][k];
            }
        }
        f->nE = nE;
        f->nslots = nslots;
        f->n_s = std::min(nE, fem_max_smem_edges(c->V));
        std::vector<int> attach_of(c->V, -1), surf_of(c->V, -1);
        for (int k = 0; k < c->A; ++k) {
            if (attach[k] < 0 || attach[k] >= c->V) { tx_fem_destroy(f); return ffail(nullptr, TX_ERR_INVALID_ARG, "tx_fem_create: attach index out of range"); }
            attach_of[attach[k]] = k;
        }
        for (int k = 0; k < c->S; ++k) {
            if (surf[k] < 0 || surf[k] >= c->V) { tx_fem_destroy(f); return ffail(nullptr, TX_ERR_INVALID_ARG, "tx_fem_create: surface index out of range"); }
            surf_of[surf[k]] = k;
        }
        std::vector<int> es;
        {
            const int TCH = f->chunk, nch = (c->T + TCH - 1) / TCH;
            es.assign((size_t)(nch + 1) * nE, 0);
            for (int e = 0; e < nE; ++e) {
                int q = eoff[e];
                for (int ck = 0; ck <= nch; ++ck) {
                    while (q < eoff[e + 1] && (eadj[q] >> 4) < ck * TCH) ++q;
                    es[(size_t)ck * nE + e] = q;
                }
            }
        }
        FEM_CUDA_C(cudaMalloc(&f->d_edge_off, sizeof(int) * es.size()));
        FEM_CUDA_C(cudaMalloc(&f->d_edge_adj, sizeof(int) * eadj.size()));
        FEM_CUDA_C(cudaMalloc(&f->d_ell, sizeof(int) * ell.size()));
        FEM_CUDA_C(cudaMalloc(&f->d_attach_of, sizeof(int) * c->V));
        FEM_CUDA_C(cudaMalloc(&f->d_surf_of, sizeof(int) * c->V));
        FEM_CUDA_C(cudaMalloc(&f->d_valg, sizeof(double) * 9 * (size_t)std::max(nE - f->n_s, 1) * f->grid));
        FEM_CUDA_C(cudaMemcpy(f->d_edge_off, es.data(), sizeof(int) * es.size(), cudaMemcpyHostToDevice));
        FEM_CUDA_C(cudaMemcpy(f->d_edge_adj, eadj.data(), sizeof(int) * eadj.size(), cudaMemcpyHostToDevice));
        FEM_CUDA_C(cudaMemcpy(f->d_ell, ell.data(), sizeof(int) * ell.size(), cudaMemcpyHostToDevice));
        FEM_CUDA_C(cudaMemcpy(f->d_attach_of, attach_of.data(), sizeof(int) * c->V, cudaMemcpyHostToDevice));
        FEM_CUDA_C(cudaMemcpy(f->d_surf_of, surf_of.data(), sizeof(int) * c->V, cudaMemcpyHostToDevice));
    }
    {   // vertex -> (tet, local vertex) incidence in CSR form, ascending tet order
        std::vector<int> off(c->V + 1, 0), adj((size_t)4 * c->T);
        for (int t = 0; t < c->T; ++t)
            for (int k = 0; k < 4; ++k) off[tets[4 * t + k] + 1]++;
        for (int i = 0; i < c->V; ++i) off[i + 1] += off[i];
        std::vector<int> cur(off.begin(), off.end() - 1);
        for (int t = 0; t < c->T; ++t)
            for (int k = 0; k < 4; ++k) adj[cur[tets[4 * t + k]]++] = 4 * t + k;
        FEM_CUDA_C(cudaMalloc(&f->d_adj_off, sizeof(int) * (c->V + 1)));
        FEM_CUDA_C(cudaMalloc(&f->d_adj, sizeof(int) * 4 * c->T));
        {   // per chunk of one-tet-per-thread: where the entries of every row start
            const int TH = fem_threads(), TCH = f->chunk, nch = (c->T + TCH - 1) / TCH;
            std::vector<int> rs((size_t)(nch + 1) * TH, 0);
            for (int i = 0; i < c->V; ++i) {
                int q = off[i];
                for (int ck = 0; ck <= nch; ++ck) {
                    while (q < off[i + 1] && (adj[q] >> 2) < ck * TCH) ++q;
                    rs[(size_t)ck * TH + i] = q;
                }
            }
            FEM_CUDA_C(cudaFree(f->d_adj_off));
            FEM_CUDA_C(cudaMalloc(&f->d_adj_off, sizeof(int) * rs.size()));
            FEM_CUDA_C(cudaMemcpy(f->d_adj_off, rs.data(), sizeof(int) * rs.size(), cudaMemcpyHostToDevice));
        }
        FEM_CUDA_C(cudaMemcpy(f->d_adj, adj.data(), sizeof(int) * 4 * c->T, cudaMemcpyHostToDevice));
    }
    FEM_CUDA_C(cudaMemcpy(f->d_tets, tets, sizeof(int) * 4 * c->T, cudaMemcpyHostToDevice));
    if (c->A > 0) FEM_CUDA_C(cudaMemcpy(f->d_attach, attach, sizeof(int) * c->A, cudaMemcpyHostToDevice));
    if (c->S > 0) FEM_CUDA_C(cudaMemcpy(f->d_surf, surf, sizeof(int) * c->S, cudaMemcpyHostToDevice));
    {   // structure-of-arrays [9][T] for coalesced per-tet loads
        std::vector<double> soa((size_t)9 * c->T);
        for (int t = 0; t < c->T; ++t)
            for (int k = 0; k < 9; ++k) soa[(size_t)k * c->T + t] = Dmi[(size_t)9 * t + k];
        FEM_CUDA_C(cudaMemcpy(f->d_Dm_inv, soa.data(), sizeof(double) * 9 * c->T, cudaMemcpyHostToDevice));
    }
    FEM_CUDA_C(cudaMemcpy(f->d_vol, vol.data(), sizeof(double) * c->T, cudaMemcpyHostToDevice));
    FEM_CUDA_C(cudaMemcpy(f->d_mass, f->mass.data(), sizeof(double) * c->V, cudaMemcpyHostToDevice));
    FEM_CUDA_C(cudaMemcpy(f->d_X, X, sizeof(double) * 3 * c->V, cudaMemcpyHostToDevice));
#undef FEM_CUDA_C
    *out = f;
    return TX_OK;
}

extern "C" void tx_fem_destroy(tx_fem* f)
{
    if (!f) return;
    cudaSetDevice(f->device);
    cudaFree(f->d_tets); cudaFree(f->d_attach); cudaFree(f->d_surf); cudaFree(f->d_Dm_inv); cudaFree(f->d_vol);
    cudaFree(f->d_mass); cudaFree(f->d_X); cudaFree(f->d_tsc); cudaFree(f->d_valg); cudaFree(f->d_xt); cudaFree(f->d_edge_off); cudaFree(f->d_edge_adj);
    cudaFree(f->d_ell); cudaFree(f->d_attach_of); cudaFree(f->d_surf_of); cudaFree(f->d_adj_off); cudaFree(f->d_adj);
    cudaFree(f->d_tri); cudaFree(f->d_w); cudaFree(f->d_top); cudaFree(f->d_mesh_tri); cudaFree(f->d_mesh_box); cudaFree(f->d_mesh_vert);
    for (int* g : f->d_grid_buf) cudaFree(g);
    cudaFree(f->d_ctri); cudaFree(f->d_ctri_row_start); cudaFree(f->d_ctri_row_adj); cudaFree(f->d_ctri_edge_start); cudaFree(f->d_ctri_edge_adj);
    cudaFree(f->d_cedge); cudaFree(f->d_cedge_row_start); cudaFree(f->d_cedge_row_adj); cudaFree(f->d_edge_cedge); cudaFree(f->d_mesh_edge); cudaFree(f->d_cedge_len2);
    delete f;
}

extern "C" int tx_fem_get_mass(const tx_fem* f, double* mass)
{
    if (!f || !mass) return TX_ERR_INVALID_ARG;
    for (int i = 0; i < f->cfg.V; ++i) mass[i] = f->mass[i];
    return TX_OK;
}

extern "C" int tx_fem_step(tx_fem* f, double* x, double* v, double* x_prev, const double* aim, const tx_fem_indenter* ind_prev,
                           const tx_fem_indenter* ind_next, int N, tx_fem_stats* stats)
{
    if (!f || !x || !v || !x_prev || !ind_prev || !ind_next || N < 0 || (f->cfg.A > 0 && !aim))
        return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_step: bad argument");
    if (N == 0) return TX_OK;
    FEM_CUDA(f, cudaSetDevice(f->device));
    const tx_fem_config& c = f->cfg;
    FemArgs a{};
    a.N = N; a.V = c.V; a.T = c.T; a.A = c.A; a.S = c.S;
    a.tets = f->d_tets; a.Dm_inv = f->d_Dm_inv; a.vol = f->d_vol; a.mass = f->d_mass; a.attach = f->d_attach; a.surf = f->d_surf;
    a.x = x; a.v = v; a.x_prev = x_prev; a.aim = aim; a.ind_prev = ind_prev; a.ind_next = ind_next; a.stats = stats;
    a.tet_scratch = f->d_tsc;
    a.val_scratch = f->d_valg;
    a.xt_scratch = f->d_xt;
    a.nE = f->nE; a.n_s = f->n_s; a.nslots = f->nslots; a.chunk = f->chunk;
    a.edge_start = f->d_edge_off; a.edge_adj = f->d_edge_adj; a.ell = f->d_ell;
    a.attach_of = f->d_attach_of; a.surf_of = f->d_surf_of;
    a.mesh_tri = f->d_mesh_tri; a.mesh_box = f->d_mesh_box; a.mesh_n = f->mesh_n;
    a.mesh_vert = f->d_mesh_vert; a.mesh_nv = f->mesh_nv;
    a.ctri = f->d_ctri; a.n_ctri = f->mesh_n > 0 ? f->n_ctri : 0;
    a.ctri_row_start = f->d_ctri_row_start; a.ctri_row_adj = f->d_ctri_row_adj;
    a.ctri_edge_start = f->d_ctri_edge_start; a.ctri_edge_adj = f->d_ctri_edge_adj;
    a.cedge = f->d_cedge; a.cedge_len2 = f->d_cedge_len2; a.n_cedge = f->mesh_n > 0 ? f->n_cedge : 0;
    a.cedge_row_start = f->d_cedge_row_start; a.cedge_row_adj = f->d_cedge_row_adj; a.edge_cedge = f->d_edge_cedge;
    a.mesh_edge = f->d_mesh_edge; a.mesh_ne = f->mesh_ne;
    a.grid_tri = f->mgrid[0]; a.grid_vert = f->mgrid[1]; a.grid_edge = f->mgrid[2];
    a.dbg_cycles = f->d_cycles;
    a.dbg_mode = getenv("TX_FEM_DBG_MODE") ? atoi(getenv("TX_FEM_DBG_MODE")) : 0;
    a.row_start = f->d_adj_off;
    a.adj = f->d_adj;
    a.dt = c.dt;
    for (int i = 0; i < 3; ++i) a.gravity[i] = c.gravity[i];
    a.mu = c.mu; a.lambda = c.lambda; a.attach_strength = c.attach_strength; a.d_hat = c.d_hat; a.kappa = c.kappa;
    a.friction_mu = c.friction_mu; a.eps_velocity = c.eps_velocity;
    a.velocity_tol = c.velocity_tol; a.pcg_tol_rate = c.pcg_tol_rate; a.newton_max_iter = c.newton_max_iter;
    a.pcg_max_iter_ratio = c.pcg_max_iter_ratio; a.ls_max_iter = c.ls_max_iter; a.substep = c.substep;
    const int grid = N < f->grid ? N : f->grid;
    FEM_CUDA(f, launch_fem_step(a, grid, f->stream));
    return TX_OK;
}

// Uniform grid over reference points `ref` [n][3] (see MeshGrid in tx_kernels.h): every primitive in the cell of its reference point,
// ids ascending within a cell. Up to 256 primitives: one cell (the brute-force loop). Cell size ~ the search radius 2 d_hat.
static cudaError_t build_mesh_grid(const std::vector<double>& ref, double rmax, double d_hat, MeshGrid& G, int*& d_start, int*& d_ids)
{
    const int n = (int)(ref.size() / 3);
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int a = 0; a < 3; ++a) {
        lo[a] = 1e300; hi[a] = -1e300;
        for (int i = 0; i < n; ++i) { lo[a] = std::min(lo[a], ref[3 * i + a]); hi[a] = std::max(hi[a], ref[3 * i + a]); }
        if (n == 0) lo[a] = hi[a] = 0.0;
    }
    int dims[3] = {1, 1, 1};
    if (n > 256) { // cells of about the search radius, but not so many that most stay empty (>= ~4 primitives per cell on average)
        const int cap = std::max(1, std::min(16, (int)std::cbrt(n / 4.0)));
        for (int a = 0; a < 3; ++a) {
            const double ext = hi[a] - lo[a];
            dims[a] = ext > 0.0 ? std::max(1, std::min(cap, (int)(ext / (2.0 * d_hat)))) : 1;
        }
    }
    G.nx = dims[0]; G.ny = dims[1]; G.nz = dims[2];
    G.rmax = rmax;
    for (int a = 0; a < 3; ++a) {
        G.lo[a] = lo[a];
        G.inv[a] = (dims[a] > 1) ? dims[a] / (hi[a] - lo[a]) : 0.0;
    }
    const int nc = dims[0] * dims[1] * dims[2];
    std::vector<int> start(nc + 1, 0), ids(std::max(n, 1)), cell(n);
    for (int i = 0; i < n; ++i) {
        int c[3];
        for (int a = 0; a < 3; ++a) {
            int k = (int)std::floor((ref[3 * i + a] - G.lo[a]) * G.inv[a]);
            c[a] = k < 0 ? 0 : (k >= dims[a] ? dims[a] - 1 : k);
        }
        cell[i] = (c[2] * dims[1] + c[1]) * dims[0] + c[0];
        start[cell[i] + 1]++;
    }
    for (int k = 0; k < nc; ++k) start[k + 1] += start[k];
    std::vector<int> cur(start.begin(), start.end() - 1);
    for (int i = 0; i < n; ++i) ids[cur[cell[i]]++] = i;
    cudaFree(d_start); cudaFree(d_ids);
    d_start = d_ids = nullptr;
    cudaError_t e = cudaMalloc(&d_start, sizeof(int) * start.size());
    if (e == cudaSuccess) e = cudaMalloc(&d_ids, sizeof(int) * ids.size());
    if (e == cudaSuccess) e = cudaMemcpy(d_start, start.data(), sizeof(int) * start.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_ids, ids.data(), sizeof(int) * ids.size(), cudaMemcpyHostToDevice);
    G.start = d_start; G.ids = d_ids;
    return e;
}

extern "C" int tx_fem_set_indenter_mesh(tx_fem* f, int n_tris, const double* tri_local)
{
    if (!f || n_tris < 0 || (n_tris > 0 && !tri_local)) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_set_indenter_mesh: bad argument");
    if (n_tris > 4096) return ffail(f, TX_ERR_UNSUPPORTED, "tx_fem_set_indenter_mesh: more than 4096 triangles (every triangle is box-tested per vertex)");
    FEM_CUDA(f, cudaSetDevice(f->device));
    FEM_CUDA(f, cudaStreamSynchronize(f->stream)); // a step in flight may still read the old mesh
    cudaFree(f->d_mesh_tri); cudaFree(f->d_mesh_box); cudaFree(f->d_mesh_vert);
    f->d_mesh_tri = f->d_mesh_box = f->d_mesh_vert = nullptr;
    f->mesh_n = f->mesh_nv = f->mesh_ne = 0;
    if (n_tris == 0) return TX_OK;
    std::vector<double> box((size_t)6 * n_tris);
    for (int t = 0; t < n_tris; ++t) {
        const double* tr = tri_local + (size_t)9 * t;
        const double e1[3] = {tr[3] - tr[0], tr[4] - tr[1], tr[5] - tr[2]}, e2[3] = {tr[6] - tr[0], tr[7] - tr[1], tr[8] - tr[2]};
        const double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
        if (!(nx * nx + ny * ny + nz * nz > 0.0)) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_set_indenter_mesh: degenerate triangle");
        for (int a = 0; a < 3; ++a) {
            double lo = tr[a], hi = tr[a];
            for (int v = 1; v < 3; ++v) {
                const double c = tr[3 * v + a];
                lo = c < lo ? c : lo;
                hi = c > hi ? c : hi;
            }
            box[(size_t)6 * t + a] = lo;
            box[(size_t)6 * t + 3 + a] = hi;
        }
    }
    FEM_CUDA(f, cudaMalloc(&f->d_mesh_tri, sizeof(double) * 9 * n_tris));
    FEM_CUDA(f, cudaMalloc(&f->d_mesh_box, sizeof(double) * 6 * n_tris));
    FEM_CUDA(f, cudaMemcpy(f->d_mesh_tri, tri_local, sizeof(double) * 9 * n_tris, cudaMemcpyHostToDevice));
    FEM_CUDA(f, cudaMemcpy(f->d_mesh_box, box.data(), sizeof(double) * 6 * n_tris, cudaMemcpyHostToDevice));
    {   // unique vertices (exact coordinate match, first occurrence order) for the vertex-vs-gel-triangle candidates
        std::vector<double> vert;
        for (int k = 0; k < 3 * n_tris; ++k) {
            const double* v = tri_local + (size_t)3 * k;
            bool found = false;
            for (size_t j = 0; j < vert.size() / 3 && !found; ++j) found = vert[3 * j] == v[0] && vert[3 * j + 1] == v[1] && vert[3 * j + 2] == v[2];
            if (!found) vert.insert(vert.end(), v, v + 3);
        }
        FEM_CUDA(f, cudaMalloc(&f->d_mesh_vert, sizeof(double) * vert.size()));
        FEM_CUDA(f, cudaMemcpy(f->d_mesh_vert, vert.data(), sizeof(double) * vert.size(), cudaMemcpyHostToDevice));
        f->mesh_nv = (int)(vert.size() / 3);
        // unique edges by vertex id, first occurrence order (edge-edge candidates)
        std::vector<int> me;
        for (int t = 0; t < n_tris; ++t) {
            int id[3];
            for (int k = 0; k < 3; ++k) {
                const double* v = tri_local + (size_t)9 * t + 3 * k;
                id[k] = 0;
                while (!(vert[3 * id[k]] == v[0] && vert[3 * id[k] + 1] == v[1] && vert[3 * id[k] + 2] == v[2])) ++id[k];
            }
            for (int k = 0; k < 3; ++k) {
                const int a = std::min(id[k], id[(k + 1) % 3]), b = std::max(id[k], id[(k + 1) % 3]);
                bool found = false;
                for (size_t j = 0; j < me.size() / 2 && !found; ++j) found = me[2 * j] == a && me[2 * j + 1] == b;
                if (!found) { me.push_back(a); me.push_back(b); }
            }
        }
        cudaFree(f->d_mesh_edge);
        f->d_mesh_edge = nullptr;
        FEM_CUDA(f, cudaMalloc(&f->d_mesh_edge, sizeof(int) * std::max<size_t>(me.size(), 1)));
        FEM_CUDA(f, cudaMemcpy(f->d_mesh_edge, me.data(), sizeof(int) * me.size(), cudaMemcpyHostToDevice));
        f->mesh_ne = (int)(me.size() / 2);
        // broad phase: reference points + largest reach of a primitive from its reference point
        std::vector<double> rt((size_t)3 * n_tris), re((size_t)3 * f->mesh_ne);
        double rmax_t = 0.0, rmax_e = 0.0;
        for (int t = 0; t < n_tris; ++t) {
            const double* tr = tri_local + (size_t)9 * t;
            for (int a = 0; a < 3; ++a) rt[3 * t + a] = (tr[a] + tr[3 + a] + tr[6 + a]) / 3.0;
            for (int k = 0; k < 3; ++k) {
                double d2 = 0.0;
                for (int a = 0; a < 3; ++a) d2 += (tr[3 * k + a] - rt[3 * t + a]) * (tr[3 * k + a] - rt[3 * t + a]);
                rmax_t = std::max(rmax_t, std::sqrt(d2));
            }
        }
        for (int q = 0; q < f->mesh_ne; ++q) {
            const double *v0 = &vert[3 * me[2 * q]], *v1 = &vert[3 * me[2 * q + 1]];
            double d2 = 0.0;
            for (int a = 0; a < 3; ++a) { re[3 * q + a] = 0.5 * (v0[a] + v1[a]); d2 += (v1[a] - v0[a]) * (v1[a] - v0[a]); }
            rmax_e = std::max(rmax_e, 0.5 * std::sqrt(d2));
        }
        FEM_CUDA(f, build_mesh_grid(rt, rmax_t, f->cfg.d_hat, f->mgrid[0], f->d_grid_buf[0], f->d_grid_buf[1]));
        FEM_CUDA(f, build_mesh_grid(vert, 0.0, f->cfg.d_hat, f->mgrid[1], f->d_grid_buf[2], f->d_grid_buf[3]));
        FEM_CUDA(f, build_mesh_grid(re, rmax_e, f->cfg.d_hat, f->mgrid[2], f->d_grid_buf[4], f->d_grid_buf[5]));
    }
    f->mesh_n = n_tris;
    return TX_OK;
}

extern "C" int tx_fem_set_contact_surface(tx_fem* f, int n_tris, const int32_t* tris)
{
    if (!f || n_tris < 0 || (n_tris > 0 && !tris)) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_set_contact_surface: bad argument");
    if (n_tris > fem_threads()) return ffail(f, TX_ERR_UNSUPPORTED, "tx_fem_set_contact_surface: more triangles than threads (one triangle per thread)");
    FEM_CUDA(f, cudaSetDevice(f->device));
    FEM_CUDA(f, cudaStreamSynchronize(f->stream));
    cudaFree(f->d_ctri); cudaFree(f->d_ctri_row_start); cudaFree(f->d_ctri_row_adj); cudaFree(f->d_ctri_edge_start); cudaFree(f->d_ctri_edge_adj);
    cudaFree(f->d_cedge); cudaFree(f->d_cedge_row_start); cudaFree(f->d_cedge_row_adj); cudaFree(f->d_edge_cedge); cudaFree(f->d_cedge_len2);
    f->d_ctri = f->d_ctri_row_start = f->d_ctri_row_adj = f->d_ctri_edge_start = f->d_ctri_edge_adj = nullptr;
    f->d_cedge = f->d_cedge_row_start = f->d_cedge_row_adj = f->d_edge_cedge = nullptr;
    f->d_cedge_len2 = nullptr;
    f->n_ctri = f->n_cedge = 0;
    if (n_tris == 0) return TX_OK;
    const int V = f->cfg.V, nE = f->nE;
    static const int PR[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    std::vector<std::vector<int>> rows(V), edges(nE);
    for (int t = 0; t < n_tris; ++t) {
        for (int j = 0; j < 3; ++j) {
            const int v = tris[3 * t + j];
            if (v < 0 || v >= V) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_set_contact_surface: vertex index out of range");
            rows[v].push_back(t << 2 | j);
        }
        for (int pr = 0; pr < 3; ++pr) {
            const int a = tris[3 * t + PR[pr][0]], b = tris[3 * t + PR[pr][1]];
            const auto it = f->edge_of.find({a < b ? a : b, a < b ? b : a});
            if (a == b || it == f->edge_of.end())
                return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_set_contact_surface: every triangle edge must be an edge of the tet mesh");
            edges[it->second].push_back(t << 2 | pr);
        }
    }
    std::vector<int> rs(V + 1, 0), ra, es(nE + 1, 0), ea;
    for (int i = 0; i < V; ++i) { ra.insert(ra.end(), rows[i].begin(), rows[i].end()); rs[i + 1] = (int)ra.size(); }
    for (int e = 0; e < nE; ++e) { ea.insert(ea.end(), edges[e].begin(), edges[e].end()); es[e + 1] = (int)ea.size(); }
    FEM_CUDA(f, cudaMalloc(&f->d_ctri, sizeof(int) * 3 * n_tris));
    FEM_CUDA(f, cudaMalloc(&f->d_ctri_row_start, sizeof(int) * rs.size()));
    FEM_CUDA(f, cudaMalloc(&f->d_ctri_row_adj, sizeof(int) * std::max<size_t>(ra.size(), 1)));
    FEM_CUDA(f, cudaMalloc(&f->d_ctri_edge_start, sizeof(int) * es.size()));
    FEM_CUDA(f, cudaMalloc(&f->d_ctri_edge_adj, sizeof(int) * std::max<size_t>(ea.size(), 1)));
    FEM_CUDA(f, cudaMemcpy(f->d_ctri, tris, sizeof(int) * 3 * n_tris, cudaMemcpyHostToDevice));
    FEM_CUDA(f, cudaMemcpy(f->d_ctri_row_start, rs.data(), sizeof(int) * rs.size(), cudaMemcpyHostToDevice));
    FEM_CUDA(f, cudaMemcpy(f->d_ctri_row_adj, ra.data(), sizeof(int) * ra.size(), cudaMemcpyHostToDevice));
    FEM_CUDA(f, cudaMemcpy(f->d_ctri_edge_start, es.data(), sizeof(int) * es.size(), cudaMemcpyHostToDevice));
    FEM_CUDA(f, cudaMemcpy(f->d_ctri_edge_adj, ea.data(), sizeof(int) * ea.size(), cudaMemcpyHostToDevice));
    {   // unique edges of the triangles (first occurrence order over the pairs (0,1) (0,2) (1,2)): one edge-edge thread each
        std::vector<int> ce, ec(nE, -1);
        std::vector<double> len2;
        for (int t = 0; t < n_tris; ++t)
            for (int pr = 0; pr < 3; ++pr) {
                const int p0 = tris[3 * t + PR[pr][0]], p1 = tris[3 * t + PR[pr][1]];
                const int a = std::min(p0, p1), b = std::max(p0, p1);
                const int e = f->edge_of.find({a, b})->second;
                if (ec[e] >= 0) continue;
                ec[e] = (int)(ce.size() / 2);
                ce.push_back(a); ce.push_back(b);
                double l2 = 0.0;
                for (int k = 0; k < 3; ++k) l2 += (f->h_X[3 * a + k] - f->h_X[3 * b + k]) * (f->h_X[3 * a + k] - f->h_X[3 * b + k]);
                len2.push_back(l2);
            }
        const int nce = (int)(ce.size() / 2);
        if (nce > fem_threads()) return ffail(f, TX_ERR_UNSUPPORTED, "tx_fem_set_contact_surface: more contact edges than threads (one edge per thread)");
        std::vector<std::vector<int>> er(V);
        for (int q = 0; q < nce; ++q) { er[ce[2 * q]].push_back(q << 1); er[ce[2 * q + 1]].push_back(q << 1 | 1); }
        std::vector<int> ers(V + 1, 0), era;
        for (int i = 0; i < V; ++i) { era.insert(era.end(), er[i].begin(), er[i].end()); ers[i + 1] = (int)era.size(); }
        FEM_CUDA(f, cudaMalloc(&f->d_cedge, sizeof(int) * ce.size()));
        FEM_CUDA(f, cudaMalloc(&f->d_cedge_len2, sizeof(double) * len2.size()));
        FEM_CUDA(f, cudaMalloc(&f->d_cedge_row_start, sizeof(int) * ers.size()));
        FEM_CUDA(f, cudaMalloc(&f->d_cedge_row_adj, sizeof(int) * era.size()));
        FEM_CUDA(f, cudaMalloc(&f->d_edge_cedge, sizeof(int) * nE));
        FEM_CUDA(f, cudaMemcpy(f->d_cedge, ce.data(), sizeof(int) * ce.size(), cudaMemcpyHostToDevice));
        FEM_CUDA(f, cudaMemcpy(f->d_cedge_len2, len2.data(), sizeof(double) * len2.size(), cudaMemcpyHostToDevice));
        FEM_CUDA(f, cudaMemcpy(f->d_cedge_row_start, ers.data(), sizeof(int) * ers.size(), cudaMemcpyHostToDevice));
        FEM_CUDA(f, cudaMemcpy(f->d_cedge_row_adj, era.data(), sizeof(int) * era.size(), cudaMemcpyHostToDevice));
        FEM_CUDA(f, cudaMemcpy(f->d_edge_cedge, ec.data(), sizeof(int) * nE, cudaMemcpyHostToDevice));
        f->n_cedge = nce;
    }
    f->n_ctri = n_tris;
    return TX_OK;
}

extern "C" int tx_fem_debug_set_cycles(tx_fem* f, long long* cycles)
{
    if (!f) return TX_ERR_INVALID_ARG;
    f->d_cycles = cycles;
    return TX_OK;
}

extern "C" int tx_fem_set_markers(tx_fem* f, int M, const int32_t* tri, const double* weights, const double* cam_R,
                                  const double* cam_t, double fx, double fy, double cx, double cy)
{
    if (!f || M <= 0 || !tri || !weights || !cam_R || !cam_t) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_set_markers: bad argument");
    FEM_CUDA(f, cudaSetDevice(f->device));
    cudaFree(f->d_tri); cudaFree(f->d_w);
    f->d_tri = nullptr; f->d_w = nullptr;
    FEM_CUDA(f, cudaMalloc(&f->d_tri, sizeof(int) * 3 * M));
    FEM_CUDA(f, cudaMalloc(&f->d_w, sizeof(double) * 3 * M));
    FEM_CUDA(f, cudaMemcpy(f->d_tri, tri, sizeof(int) * 3 * M, cudaMemcpyHostToDevice));
    FEM_CUDA(f, cudaMemcpy(f->d_w, weights, sizeof(double) * 3 * M, cudaMemcpyHostToDevice));
    f->M = M;
    for (int i = 0; i < 9; ++i) f->cam_R[i] = cam_R[i];
    for (int i = 0; i < 3; ++i) f->cam_t[i] = cam_t[i];
    f->fx = fx; f->fy = fy; f->cx = cx; f->cy = cy;
    return TX_OK;
}

extern "C" int tx_fem_attachment_aim(tx_fem* f, const float* pose, const float* offsets, int N, int per_env_offsets, double* aim)
{
    if (!f || !pose || !offsets || !aim || N < 0) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_attachment_aim: bad argument");
    if (f->cfg.A <= 0) return ffail(f, TX_ERR_STATE, "tx_fem_attachment_aim: the mesh has no attached vertices");
    if (N == 0) return TX_OK;
    FEM_CUDA(f, cudaSetDevice(f->device));
    FEM_CUDA(f, launch_attachment_aim(pose, offsets, N, f->cfg.A, per_env_offsets ? 1 : 0, aim, f->stream));
    return TX_OK;
}

extern "C" int tx_fem_set_surface(tx_fem* f, int n_tris, const int32_t* tris)
{
    if (!f || n_tris <= 0 || !tris) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_set_surface: bad argument");
    for (int i = 0; i < 3 * n_tris; ++i)
        if (tris[i] < 0 || tris[i] >= f->cfg.V) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_set_surface: vertex index out of range");
    FEM_CUDA(f, cudaSetDevice(f->device));
    cudaFree(f->d_top);
    f->d_top = nullptr;
    FEM_CUDA(f, cudaMalloc(&f->d_top, sizeof(int) * 3 * n_tris));
    FEM_CUDA(f, cudaMemcpy(f->d_top, tris, sizeof(int) * 3 * n_tris, cudaMemcpyHostToDevice));
    f->n_top = n_tris;
    return TX_OK;
}

extern "C" int tx_fem_heightmap(tx_fem* f, const double* x, int N, float* height_mm, int H, int W, double pitch_m, double origin_x,
                                double origin_y, double cam_z_m, float far_mm)
{
    if (!f || !x || !height_mm || N < 0 || H <= 0 || W <= 0 || !(pitch_m > 0.0) || !(far_mm > 0.0f))
        return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_heightmap: bad argument");
    if (f->n_top <= 0) return ffail(f, TX_ERR_STATE, "tx_fem_heightmap: call tx_fem_set_surface first");
    if (N == 0) return TX_OK;
    FEM_CUDA(f, cudaSetDevice(f->device));
    RasterArgs a{};
    a.x = x; a.tris = f->d_top; a.hm = height_mm; a.V = f->cfg.V; a.n_tris = f->n_top; a.H = H; a.W = W;
    a.pitch = pitch_m; a.ox = origin_x; a.oy = origin_y; a.cam_z = cam_z_m; a.far_mm = far_mm;
    FEM_CUDA(f, launch_heightmap(a, N, f->stream));
    return TX_OK;
}

extern "C" int tx_fem_set_marker_output(tx_fem* f, int normalize, double img_w, int zero_all)
{
    if (!f || !(img_w > 0.0)) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_set_marker_output: bad argument");
    f->normalize = normalize ? 1 : 0;
    f->half_w = img_w / 2.0;
    f->zero_all = zero_all ? 1 : 0;
    return TX_OK;
}

extern "C" int tx_fem_markers(tx_fem* f, const double* x, int N, float* markers)
{
    if (!f || !x || !markers || N < 0) return ffail(f, TX_ERR_INVALID_ARG, "tx_fem_markers: bad argument");
    if (f->M <= 0) return ffail(f, TX_ERR_STATE, "tx_fem_markers: call tx_fem_set_markers first");
    if (N == 0) return TX_OK;
    FEM_CUDA(f, cudaSetDevice(f->device));
    FemMarkerArgs m{};
    m.V = f->cfg.V; m.M = f->M; m.tri = f->d_tri; m.weights = f->d_w; m.x_rest = f->d_X; m.x = x; m.out = markers;
    for (int i = 0; i < 9; ++i) m.cam_R[i] = f->cam_R[i];
    for (int i = 0; i < 3; ++i) m.cam_t[i] = f->cam_t[i];
    m.fx = f->fx; m.fy = f->fy; m.cx = f->cx; m.cy = f->cy;
    m.normalize = f->normalize; m.half_w = f->half_w; m.zero_all = f->zero_all;
    FEM_CUDA(f, launch_fem_markers(m, N, f->stream));
    return TX_OK;
}
