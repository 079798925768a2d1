// C ABI of the B200-native force / energy / integrator path (include/npsl.h).
//
// Host side of the kernels in pair_kernel.cuh and mesh_kernels.cuh: flattens the static mesh
// topology once (the reference re-walks pointer-linked VERTEX/EDGE/FACE objects every call,
// src/newforces.cpp:149-265), owns all device memory, and issues one CUDA-graph launch per MD
// step. Row-sharded contexts (one process per GPU) follow the reference's MPI row decomposition
// (src/main.cpp:448-455) with an NCCL all-gather of positions and of the per-rank kinetic
// energies; NCCL is bound at run time with dlopen so that single-GPU users need no libnccl.
//
// There is no CPU path in this file: without a CUDA device every entry point fails.
#include "../../include/npsl.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "ion_kernels.cuh"
#include "mesh_kernels.cuh"
#include "pair_kernel.cuh"
#include "pair_sym_kernel.cuh"

using namespace npsl;

// ------------------------------------------------------------------------------------------------
// NCCL, bound lazily (prototypes restated from the public nccl.h ABI; stable across 2.x)
namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[NPSL_NCCL_ID_BYTES]; } ncclUniqueId;
enum { ncclFloat64_ = 8 };
struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::string g_error;  // creation-time failures (no context yet)

bool load_nccl(std::string &err) {
    if (g_nccl.ok) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) {
        err = std::string("cannot load libnccl: ") + dlerror();
        return false;
    }
#define BIND(field, sym)                                                        \
    *(void **) (&g_nccl.field) = dlsym(g_nccl.handle, sym);                     \
    if (!g_nccl.field) { err = std::string("libnccl lacks ") + sym; return false; }
    BIND(GetUniqueId, "ncclGetUniqueId")
    BIND(CommInitRank, "ncclCommInitRank")
    BIND(CommDestroy, "ncclCommDestroy")
    BIND(AllGather, "ncclAllGather")
    BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
    g_nccl.ok = true;
    return true;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
struct npsl_ctx {
    int device = 0;
    int nv = 0, ne = 0, nf = 0;
    int rank = 0, world = 1;
    int row_lo = 0, row_hi = 0;  // owned rows / vertices, half-open
    int range = 0;               // rows per rank (ceil(nv/world)); buffers hold world*range records
    npsl_params prm;
    bool all_q_zero = true;
    bool forces_valid = false;

    // state, one 32-byte record per vertex (coalesced 16-byte accesses everywhere)
    double4 *xyzq = nullptr, *vel = nullptr, *force = nullptr, *fmesh = nullptr, *ngv = nullptr;
    double *part_con = nullptr;  // [8][face blocks] partials of the SHAKE / RATTLE face sums
    double4 *comp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // pair, bend, stretch, tension, voltension
    // topology
    int4 *face_v = nullptr, *edge_rec = nullptr;
    double *l0 = nullptr;
    int *degree = nullptr, *ring = nullptr, *bslot_ref = nullptr;
    double *ring_l0 = nullptr;
    // work buffers
    double4 *bslot = nullptr, *pair_partial = nullptr, *lj_out = nullptr;
    int *lj_hit = nullptr;
    double *scal = nullptr, *scal_all = nullptr, *ke_warp = nullptr;
    int n_ke_warps = 0;  // world * range / 32
    Thermo *th_cur = nullptr, *th_mid = nullptr;
    double *part_mesh = nullptr, *part_vert = nullptr;
    unsigned int *tickets = nullptr;
    double4 *face_dir_area = nullptr;
    double *face_vol = nullptr, *edge_S = nullptr;
    // data feeds of the writers (local energy maps, vertex areas / normals)
    double2 *edge_E = nullptr, *vertex_e = nullptr;
    int4 *face_e = nullptr;
    double4 *face_local = nullptr, *vertex_geo = nullptr;
    bool local_valid = false;  // vertex_e / edge_E hold the energy pass at the current positions
    // explicit counterions (ion_kernels.cuh)
    IonParams ip;
    double4 *ion_xyzq = nullptr, *ion_vel = nullptr, *ion_force = nullptr, *ion_mesh = nullptr;
    double2 *ion_e = nullptr;
    double *ion_mesh_lj = nullptr, *ion_part = nullptr;  // ion_part: per-CTA sums of the vertices' pull on the ions
    double *peak_out = nullptr;

    // launch geometry
    int pair_rows_per_thread = 2, pair_row_tiles = 0, pair_splits = 1, pair_cols_per_cta = 0, pair_row_stride = 0;
    int n_face_blocks = 0, n_edge_blocks = 0, n_vertex_blocks = 0;
    PairParams pp;
    LjParams ljp;
    MeshParams mp;

    cudaStream_t s_main = nullptr, s_side = nullptr, s_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_lj_fork = nullptr, ev_lj_join = nullptr;
    cudaGraphExec_t step_graph = nullptr;
    // packed host-state step (npsl_md_step_packed): pinned block [pos | vel | force | KE], its device images and a
    // graph of its own (copy in -> the step, reading / writing the packed images -> copies out)
    double *h_state = nullptr, *d_state_in = nullptr, *d_state_out = nullptr;
    cudaGraphExec_t packed_graph = nullptr;
    int packed_launches = 0;
    cudaStream_t s_copy = nullptr;
    cudaEvent_t ev_copy_fork = nullptr, ev_copy_join = nullptr;
    ncclComm_t comm = nullptr;

    // measurement
    int64_t launches = 0;
    int launches_per_step = 0;
    bool timing = false;
    std::vector<cudaEvent_t> t_ev;  // pairs of events around k_pair
    size_t t_used = 0;
    double t_sum_ms = 0;
    int64_t t_count = 0;

    // host<->device staging for the packed [n][3] arrays of the boundary: pinned host + device
    double *h_pin[3] = {nullptr, nullptr, nullptr};
    double *d_pack[3] = {nullptr, nullptr, nullptr};
    // measurement: L2 flush buffer and per-step event brackets of npsl_step_timing
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    std::vector<cudaEvent_t> s_ev;
    int plan_rows = 0, plan_splits = 0;  // overrides of npsl_set_pair_plan (0 = automatic)
    // symmetric (Newton's-third-law) pair path, pair_sym_kernel.cuh
    int pair_mode = 0;       // 0 automatic, 1 ordered (k_pair), 2 symmetric (k_pair_sym)
    int sym_rows = 6, sym_threads = 64, sym_minb = 0;  // k_pair_sym variant: rows per thread, threads per CTA
    bool use_sym = false;
    SymParams sp;
    int2 *sym_units = nullptr;   // this rank's units, heaviest first
    double *sym_rowpart = nullptr, *sym_colpart = nullptr, *sym_fv = nullptr;
    int sym_n_units = 0, sym_n_units_total = 0, sym_v_first = 0, sym_v_count = kSymV;
    int *red_groups = nullptr;  // k_pair_reduce: vertex groups of the CTAs (rows + columns first, then columns only)
    int red_n_a = 0, red_n_b_groups = 0, red_n_b_ctas = 0;
    double *pair_etab = nullptr;   // SYM_FAST: index-biased table of the exponential (pair_sym_kernel.cuh), device copy
    double2 *pair_ptab = nullptr;  // piecewise-polynomial table of the radial factor (pair_sym_kernel.cuh), device copy
    int sym_math = SYM_FAST;       // how k_pair_sym evaluates the radial factor (pair_sym_kernel.cuh); NPSL_SYM_MATH
    bool sym_shape_set = false;    // rows / threads chosen through npsl_set_pair_plan
    int sym_unroll = 1;            // pair evaluations in flight per thread = 4 x this; NPSL_SYM_UNROLL
    unsigned long long *sym_trace = nullptr;  // NPSL_SYM_TRACE
    std::vector<int2> sym_trace_units;
    int n_sm = 148;
    // caller buffers of the host-buffer entry points that were page-locked with cudaHostRegister (so that the
    // copies go straight from / to the caller's arrays); anything that cannot be registered is staged
    std::vector<std::pair<void *, size_t>> pinned_user;
    // npsl_step_profile: events recorded between the launches of a step
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<const char *> prof_names;
    size_t prof_used = 0;
    // exchanges over NVLink peer memory (xchg.cuh); NCCL all-gathers are the fallback
    bool p2p = false;
    // sharding plan of a multi-GPU context: rows (owned vertex ranges, three exchanges per step) or replicated
    // integrator (every rank integrates all vertices, only the pair term is split, one exchange per step)
    bool replicated = false;
    int shard_pref = 0;                           // 0 automatic, 1 rows, 2 replicated (NPSL_SHARD)
    Xchg xg;                                      // kernel parameter block
    unsigned long long *xflags = nullptr;         // [XK_KINDS][kMaxRanks], written by the peers
    unsigned long long *xstate = nullptr;         // epochs
    unsigned int *xtickets = nullptr;
    int *x_timed_out = nullptr;                   // set by a wait that gave up (xchg.cuh)
    int *lj_all = nullptr;                        // NCCL exchanges: every rank's collision flag
    void *peer_map[5][kMaxRanks] = {};            // host copies of the mapped peer pointers (xyzq, fv, ke, flags, packed image)
    double **d_peer_state = nullptr;
    double4 **d_peer_xyzq = nullptr;
    double **d_peer_fv = nullptr, **d_peer_ke = nullptr;
    unsigned long long **d_peer_flags = nullptr;
    std::string err;
};

namespace {

int fail(npsl_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    g_error = buf;
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(c, NPSL_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define NC(call)                                                                                          \
    do {                                                                                                  \
        int e_ = (call);                                                                                  \
        if (e_ != 0)                                                                                      \
            return fail(c, NPSL_ERR_NCCL, "%s: %s (%s:%d)", #call, g_nccl.GetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <class T>
cudaError_t dalloc(T **p, size_t n) {
    cudaError_t e = cudaMalloc((void **) p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T));
    return e;
}

inline uint64_t dkey(int a, int b) { return ((uint64_t) (uint32_t) a << 32) | (uint32_t) b; }

// Static topology tables consumed by k_mesh / k_vertex (layout described in mesh_kernels.cuh).
struct Tables {
    std::vector<int4> face_v, edge_rec, face_e;
    std::vector<int> degree, ring, bslot_ref;
    std::vector<double> ring_l0;
};

// Flattens EDGE::itsV/itsF, FACE::itsV and VERTEX::itsE/itsF (src/interface.cpp:100-143) into index
// tables. Requires a closed, consistently oriented triangle mesh with valence <= NPSL_MAX_DEGREE
// (the reference asserts two faces per edge, src/interface.cpp:146).
int build_tables(const npsl_mesh_desc *m, Tables &t, std::string &why) {
    const int nv = m->n_vertices, ne = m->n_edges, nf = m->n_faces;
    std::unordered_map<uint64_t, int> third;  // directed edge a->b  ->  vertex opposite in its face
    third.reserve((size_t) nf * 3 * 2);
    t.face_v.resize(nf);
    std::vector<int> cnt(nv, 0);
    std::vector<int> nxt((size_t) nv * kMaxDeg * 2, -1);  // per vertex: (n, nn) of each incident face
    for (int f = 0; f < nf; f++) {
        const int32_t *fv = m->face_v + 3 * f;
        for (int k = 0; k < 3; k++)
            if (fv[k] < 0 || fv[k] >= nv) { why = "face vertex index out of range"; return -1; }
        if (fv[0] == fv[1] || fv[1] == fv[2] || fv[0] == fv[2]) { why = "degenerate face"; return -1; }
        t.face_v[f] = make_int4(fv[0], fv[1], fv[2], 0);
        for (int k = 0; k < 3; k++) {
            const int a = fv[k], b = fv[(k + 1) % 3], c = fv[(k + 2) % 3];
            if (!third.emplace(dkey(a, b), c).second) { why = "inconsistent face orientation"; return -1; }
            if (cnt[a] >= kMaxDeg) { why = "vertex valence exceeds NPSL_MAX_DEGREE"; return -1; }
            nxt[((size_t) a * kMaxDeg + cnt[a]) * 2] = b;
            nxt[((size_t) a * kMaxDeg + cnt[a]) * 2 + 1] = c;
            cnt[a]++;
        }
    }
    // edges: oriented record (v0, v1, a0, a1) with F0 = (v0,v1,a0), F1 = (v1,v0,a1)
    t.edge_rec.resize(ne);
    std::unordered_map<uint64_t, int> edge_of;
    edge_of.reserve((size_t) ne * 2);
    std::vector<int> ecnt(nv, 0);
    for (int e = 0; e < ne; e++) {
        const int v0 = m->edge_v[2 * e], v1 = m->edge_v[2 * e + 1];
        if (v0 < 0 || v0 >= nv || v1 < 0 || v1 >= nv || v0 == v1) { why = "bad edge"; return -1; }
        auto f0 = third.find(dkey(v0, v1)), f1 = third.find(dkey(v1, v0));
        if (f0 == third.end() || f1 == third.end()) { why = "edge without two faces (mesh not closed)"; return -1; }
        t.edge_rec[e] = make_int4(v0, v1, f0->second, f1->second);
        if (!edge_of.emplace(dkey(std::min(v0, v1), std::max(v0, v1)), e).second) { why = "duplicate edge"; return -1; }
        ecnt[v0]++;
        ecnt[v1]++;
    }
    if ((size_t) ne * 2 != third.size()) { why = "edge / face counts inconsistent"; return -1; }
    // the three edges of every face, ascending edge index
    t.face_e.resize(nf);
    for (int f = 0; f < nf; f++) {
        const int32_t *fv = m->face_v + 3 * f;
        int es[3];
        for (int k = 0; k < 3; k++) {
            const int a = fv[k], b = fv[(k + 1) % 3];
            es[k] = edge_of.at(dkey(std::min(a, b), std::max(a, b)));
        }
        std::sort(es, es + 3);
        t.face_e[f] = make_int4(es[0], es[1], es[2], 0);
    }
    // one-rings, counter-clockwise, starting at the first incident face in file order
    t.degree.assign(nv, 0);
    t.ring.assign((size_t) kMaxDeg * nv, -1);
    t.ring_l0.assign((size_t) kMaxDeg * nv, 1.0);
    for (int v = 0; v < nv; v++) {
        const int d = cnt[v];
        if (d < 3 || ecnt[v] != d) { why = "vertex is not an interior manifold vertex"; return -1; }
        t.degree[v] = d;
        const int *nx = &nxt[(size_t) v * kMaxDeg * 2];
        int cur = nx[0];
        for (int k = 0; k < d; k++) {
            t.ring[(size_t) k * nv + v] = cur;
            auto it = edge_of.find(dkey(std::min(v, cur), std::max(v, cur)));
            if (it == edge_of.end()) { why = "face side without an edge record"; return -1; }
            t.ring_l0[(size_t) k * nv + v] = m->l0[it->second];
            int found = -1;
            for (int j = 0; j < d; j++)
                if (nx[2 * j] == cur) found = nx[2 * j + 1];
            if (found < 0) { why = "open one-ring"; return -1; }
            cur = found;
        }
        if (cur != nx[0]) { why = "one-ring does not close"; return -1; }
    }
    // bending slots gathered by each vertex, ascending edge index (the order in which the
    // reference's edge loop adds into VERTEX::bforce, src/newforces.cpp:256-257)
    t.bslot_ref.assign((size_t) 2 * kMaxDeg * nv, -1);
    std::vector<int> bc(nv, 0);
    for (int e = 0; e < ne; e++) {
        const int4 r = t.edge_rec[e];
        const int vs[4] = {r.x, r.y, r.z, r.w};
        for (int s = 0; s < 4; s++) {
            const int v = vs[s];
            if (bc[v] >= 2 * kMaxDeg) { why = "bending slot overflow"; return -1; }
            t.bslot_ref[(size_t) bc[v] * nv + v] = s * ne + e;
            bc[v]++;
        }
    }
    for (int v = 0; v < nv; v++)
        if (bc[v] != 2 * t.degree[v]) { why = "bending slot count mismatch"; return -1; }
    return 0;
}

void fill_thermo(Thermo &t, const npsl_bath_link *mesh, const npsl_bath_link *ions) {
    for (int j = 0; j < kMaxChain; j++) {
        t.mesh[j] = {mesh[j].Q, mesh[j].T, mesh[j].dof, mesh[j].xi, mesh[j].eta};
        t.ions[j] = {ions[j].Q, ions[j].T, ions[j].dof, ions[j].xi, ions[j].eta};
    }
}

int high_word(double x) {
    int64_t b;
    memcpy(&b, &x, 8);
    return (int) (b >> 32);
}

// Column chunk / row tile geometry of k_pair. The chunk width is a function of the vertex count
// only -- never of the rank count, the owned row range or the SM count -- so every (chunk,row)
// partial sum, and with it the whole trajectory, is bit-identical however the rows are sharded
// across GPUs: 256 * max(2, ceil(N / 8192)) columns, i.e. at most 32 chunks (27 x 1536 at S64,
// 31 x 5376 at S128). Rows per thread only changes which thread computes a row, not its value.
void plan_pair(npsl_ctx *c, int n_sm) {
    const int rows = c->row_hi - c->row_lo;
    const int col_tiles = (c->nv + kPairTileJ - 1) / kPairTileJ;
    int tiles_per_cta = std::max(2, (c->nv + 32 * kPairTileJ - 1) / (32 * kPairTileJ));
    if (c->plan_splits) {  // measurement override (changes the grouping of the column sum)
        const int splits = std::max(1, std::min(c->plan_splits, col_tiles));
        tiles_per_cta = (col_tiles + splits - 1) / splits;
    }
    c->pair_splits = (col_tiles + tiles_per_cta - 1) / tiles_per_cta;
    c->pair_cols_per_cta = tiles_per_cta * kPairTileJ;
    // two rows per thread unless that leaves fewer than ~4 CTAs per SM (small shards)
    int rpt = 2;
    if ((long) ((rows + 2 * kPairThreads - 1) / (2 * kPairThreads)) * c->pair_splits < 4L * n_sm) rpt = 1;
    c->pair_rows_per_thread = c->plan_rows ? c->plan_rows : rpt;
    const int tile_rows = kPairThreads * c->pair_rows_per_thread;
    c->pair_row_tiles = std::max(1, (rows + tile_rows - 1) / tile_rows);
    c->pair_row_stride = (rows + 31) / 32 * 32;
}

// copies the launch plan into the kernel parameter blocks
void apply_pair_plan(npsl_ctx *c) {
    c->pp.cols_per_cta = c->pair_cols_per_cta;
    c->pp.n_splits = c->pair_splits;
    c->pp.row_stride = c->pair_row_stride;
    c->mp.n_splits = c->pair_splits;
    c->mp.row_stride = c->pair_row_stride;
}

// Column-chunk geometry of the symmetric pair kernel: a function of the vertex count and the kernel variant's
// row-block size only (plus the two measurement knobs), never of the rank count.
void sym_geometry(int nv, int rb, SymParams &sp) {
    sp.n = nv;
    sp.npad = (nv + kSymPad - 1) / kSymPad * kSymPad;
    sp.rb = rb;
    // 256 columns per unit up to 65536 vertices, at most 256 such chunks beyond: fine-grained enough to fill 8 GPUs.
    sp.chunk_tiles = std::max(256 / kSymCT, (nv + 256 * kSymCT - 1) / (256 * kSymCT));
    if (const char *e = getenv("NPSL_SYM_CHUNK_TILES"))  // measurement knob; changes the grouping of the row sums
        if (atoi(e) > 0) sp.chunk_tiles = atoi(e);
    const int n_tiles = (nv + kSymCT - 1) / kSymCT;
    // The column range ends in narrower chunks: half width over its last 12 %, of which the last quarter an eighth of the
    // width (one 32-column group at S64, both warps of a CTA on it at once). Units are launched heaviest first, so these
    // run last and even out the ragged end of the launch: a rank of an 8-GPU job at S64 has about seven 64-column tiles
    // of work per resident CTA. The mix was chosen by measurement on one GPU playing one rank of 8 and on the full
    // launch (profiles/r02_unit_geometry.txt): a CTA that arrives beside an older CTA on its SM sub-partitions runs at
    // half speed until the older one leaves (21 against 9.8 us per 32 columns), so a mid-sized unit dispatched late is what
    // ends a rank's launch -- few mid-sized units, many small ones. A function of the vertex count only, like everything here.
    int fine_pct = 12, finest_pct = 3;
    if (const char *e = getenv("NPSL_SYM_FINE_PCT")) fine_pct = std::max(0, std::min(100, atoi(e)));      // measurement knobs
    if (const char *e = getenv("NPSL_SYM_FINEST_PCT")) finest_pct = std::max(0, std::min(100, atoi(e)));
    finest_pct = std::min(finest_pct, fine_pct);
    sp.fine_tiles = std::max(1, sp.chunk_tiles / 2);
    sp.finest_tiles = std::max(1, sp.chunk_tiles / 8);
    if (const char *e = getenv("NPSL_SYM_FINE_TILES"))
        if (atoi(e) > 0) sp.fine_tiles = atoi(e);
    if (const char *e = getenv("NPSL_SYM_FINEST_TILES"))
        if (atoi(e) > 0) sp.finest_tiles = atoi(e);
    const int coarse_tiles = n_tiles - (int) ((long) n_tiles * fine_pct / 100);
    sp.n_coarse = fine_pct >= 100 ? 0 : (coarse_tiles + sp.chunk_tiles - 1) / sp.chunk_tiles;
    const int rest = std::max(0, n_tiles - sp.n_coarse * sp.chunk_tiles);
    const int finest_rest = std::min(rest, (int) ((long) n_tiles * finest_pct / 100));
    sp.n_fine = finest_pct >= 100 ? 0 : (rest - finest_rest + sp.fine_tiles - 1) / sp.fine_tiles;
    const int rest2 = std::max(0, rest - sp.n_fine * sp.fine_tiles);
    sp.n_chunks = sp.n_coarse + sp.n_fine + (rest2 + sp.finest_tiles - 1) / sp.finest_tiles;
}

// Radial factor of the screened-Coulomb force, G(s) = e^{-r/lambda} (1/r^3 + 1/(lambda r^2)), r = sqrt(s), in
// extended precision (what src/newforces.cpp:173-174 multiplies r_vec with, up to the charge prefactor).
long double radial_factor(long double s, long double lambda) {
    const long double r = sqrtl(s);
    return expl(-r / lambda) * (1.0L / (r * r * r) + 1.0L / (lambda * r * r));
}

// Table of pair_sym_kernel.cuh: per interval [lo, lo + w) of s (kPolyBits mantissa bits), the degree-5 polynomial in
// t = s - (lo + w/2) that interpolates G at the six Chebyshev nodes of the interval; coefficients c0..c5 as doubles.
// Everything is evaluated in long double (the fit is a 6-term Chebyshev series converted to monomials in t / (w/2)).
void build_pair_table(double lambda, std::vector<double> &tab) {
    constexpr int N = 6;
    tab.assign((size_t) kPolyCount * N, 0.0);
    const long double pi = 3.14159265358979323846264338327950288L;
    // Chebyshev -> monomial: T0 = 1, T1 = x, T2 = 2x^2-1, T3 = 4x^3-3x, T4 = 8x^4-8x^2+1, T5 = 16x^5-20x^3+5x
    static const long double T[N][N] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {-1, 0, 2, 0, 0, 0},
                                        {0, -3, 0, 4, 0, 0}, {1, 0, -8, 0, 8, 0}, {0, 5, 0, -20, 0, 16}};
    for (int k = 0; k < kPolyCount; k++) {
        const int oct = k >> kPolyBits, j = k & ((1 << kPolyBits) - 1);
        const long double w = ldexpl(1.0L, kPolyEmin + oct - kPolyBits);
        const long double mid = ldexpl(1.0L + (long double) j / (1 << kPolyBits), kPolyEmin + oct) + 0.5L * w;
        const long double h = 0.5L * w;
        long double f[N], a[N];
        for (int m = 0; m < N; m++) f[m] = radial_factor(mid + h * cosl(pi * (m + 0.5L) / N), lambda);
        for (int n = 0; n < N; n++) {
            long double sum = 0;
            for (int m = 0; m < N; m++) sum += f[m] * cosl(pi * n * (m + 0.5L) / N);
            a[n] = sum * (n == 0 ? 1.0L : 2.0L) / N;
        }
        long double hp = 1.0L;
        for (int p = 0; p < N; p++) {  // coefficient of t^p
            long double b = 0;
            for (int n = 0; n < N; n++) b += a[n] * T[n][p];
            tab[(size_t) k * N + p] = (double) (b / hp);
            hp *= h;
        }
    }
}

// The polynomial of the table at s, with the device's operations (interval from the high word, Horner in FP64);
// used by npsl_pair_poly_eval so that the accuracy of the table can be checked without a GPU.
double eval_pair_table(const std::vector<double> &tab, double s) {
    int64_t bits;
    memcpy(&bits, &s, 8);
    const int hi = (int) (bits >> 32);
    const int k = std::min(std::max((hi >> kPolyShift) - kPolyBase, 0), kPolyCount - 1);
    const int64_t mid_bits = (int64_t) ((hi & (int) (0xffffffffu << kPolyShift)) | (1 << (kPolyShift - 1))) << 32;
    double mid;
    memcpy(&mid, &mid_bits, 8);
    const double t = s - mid;
    const double *c = &tab[(size_t) k * 6];
    double g = fma(c[5], t, c[4]);
    g = fma(g, t, c[3]);
    g = fma(g, t, c[2]);
    g = fma(g, t, c[1]);
    return fma(g, t, c[0]);
}

struct SymUnit { int I, cc, work; };

// Units (row block, column chunk) of the virtual ranks [v_first, v_first + v_count), heaviest first; returns the
// number of units of all virtual ranks. A row block's units all belong to sym_vrank(row block).
int sym_units_of(const SymParams &sp, int v_first, int v_count, std::vector<SymUnit> &mine) {
    const int n_tiles = (sp.n + kSymCT - 1) / kSymCT, n_rb = (sp.n + sp.rb - 1) / sp.rb;
    int total = 0;
    mine.clear();
    for (int I = 0; I < n_rb; I++) {
        const int v = sym_vrank(I);
        for (int cc = 0; cc < sp.n_chunks; cc++) {
            const int tile_end = std::min(sym_chunk_tile0(sp, cc + 1), n_tiles);
            const int tile_begin = std::max(sym_chunk_tile0(sp, cc), I * sp.rb / kSymCT);
            if (tile_end <= tile_begin) continue;
            total++;
            if (v >= v_first && v < v_first + v_count) mine.push_back({I, cc, tile_end - tile_begin});
        }
    }
    std::stable_sort(mine.begin(), mine.end(), [](const SymUnit &a, const SymUnit &b) { return a.work > b.work; });
    return total;
}

// Units of the symmetric pair kernel. Everything here depends on the vertex count only, except which of
// the 8 virtual ranks this rank computes: [rank*8/world, (rank+1)*8/world).
int plan_sym(npsl_ctx *c) {
    const int nv = c->nv;
    SymParams &sp = c->sp;
    if (const char *e = getenv("NPSL_SYM_MATH")) c->sym_math = std::max(0, std::min(2, atoi(e)));
    if (const char *e = getenv("NPSL_SYM_UNROLL")) c->sym_unroll = atoi(e) == 2 ? 2 : 1;
    if (!c->sym_shape_set && (c->sym_math != SYM_FAST || c->sym_unroll != 1)) { c->sym_rows = 4; c->sym_threads = 128; }
    sym_geometry(nv, c->sym_rows * c->sym_threads, sp);
    sp.c2s = c->pp.c2s; sp.inv_lambda = c->pp.inv_lambda; sp.lj_cut_hi = c->pp.lj_cut_hi;
    if (!c->pair_ptab) {
        std::vector<double> tab;
        build_pair_table(c->prm.inv_kappa_out, tab);
        CU(dalloc(&c->pair_ptab, tab.size() / 2));
        CU(cudaMemcpy(c->pair_ptab, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    }
    sp.ptab = c->pair_ptab;
    if (!c->pair_etab) {  // SYM_FAST: the exponential's table with the index folded into the high words (pair_sym_kernel.cuh)
        static const double tab[kExpTabFast] = {NPSL_EXP2FAST_TABLE};
        std::vector<double> biased(kExpTabFast);
        for (int k = 0; k < kExpTabFast; k++) {
            uint64_t bits;
            memcpy(&bits, &tab[k], 8);
            bits -= (uint64_t) k << (32 + 20 - NPSL_EXP2FAST_TB);  // high word - (k << (20 - TB)), no borrow into it from below
            memcpy(&biased[k], &bits, 8);
        }
        CU(dalloc(&c->pair_etab, biased.size()));
        CU(cudaMemcpy(c->pair_etab, biased.data(), sizeof(double) * biased.size(), cudaMemcpyHostToDevice));
    }
    sp.etab = c->pair_etab;
    sp.fast_c2s = 1.4426950408889634074 * (double) kExpTabFast; sp.fast_g1 = NPSL_EXP2FAST_G1; sp.fast_g2 = NPSL_EXP2FAST_G2;

    const int n_rb = (nv + sp.rb - 1) / sp.rb;
    const bool want = c->pair_mode == 2 || (c->pair_mode == 0 && nv >= 4096);
    c->use_sym = want && (kSymV % c->world == 0);
    c->mp.pair_sym = c->use_sym ? 1 : 0;
    c->mp.npad = sp.npad;
    if (!c->use_sym) return 0;
    c->sym_v_count = kSymV / c->world;
    c->sym_v_first = c->rank * c->sym_v_count;
    // Measurement only: NPSL_EMULATE_SHARD="r/w" makes a single-GPU context do just the pair work of rank r of w
    // (incomplete forces), so that a rank's kernels can be profiled under ncu without a multi-rank job.
    if (const char *e = getenv("NPSL_EMULATE_SHARD")) {
        int er = 0, ew = 1;
        if (c->world == 1 && sscanf(e, "%d/%d", &er, &ew) == 2 && ew > 0 && kSymV % ew == 0 && er >= 0 && er < ew) {
            c->sym_v_count = kSymV / ew;
            c->sym_v_first = er * c->sym_v_count;
        }
    }
    std::vector<SymUnit> mine;
    c->sym_n_units_total = sym_units_of(sp, c->sym_v_first, c->sym_v_count, mine);
    std::vector<int2> units(mine.size());
    for (size_t k = 0; k < mine.size(); k++) units[k] = make_int2(mine[k].I, mine[k].cc);
    c->sym_n_units = (int) units.size();
    void *old[] = {c->sym_units, c->sym_rowpart, c->sym_colpart};
    for (void *q : old)
        if (q) cudaFree(q);
    c->sym_units = nullptr; c->sym_rowpart = c->sym_colpart = nullptr;
    CU(dalloc(&c->sym_units, units.size()));
    CU(dalloc(&c->sym_rowpart, (size_t) sp.n_chunks * 3 * sp.npad));
    CU(dalloc(&c->sym_colpart, (size_t) n_rb * 3 * sp.npad));
    if (!c->sym_fv) CU(dalloc(&c->sym_fv, (size_t) 2 * kSymV * 3 * sp.npad));  // fixed size: peers map it; two epochs
    CU(cudaMemcpy(c->sym_units, units.data(), sizeof(int2) * units.size(), cudaMemcpyHostToDevice));
    {  // k_pair_reduce: groups of 32 vertices whose row block is computed here come first (they also add row partials)
        std::vector<int> ga, gb;
        const int n_groups = (nv + 31) / 32;
        for (int g = 0; g < n_groups; g++) {
            const int v = sym_vrank(g * 32 / sp.rb);
            (v >= c->sym_v_first && v < c->sym_v_first + c->sym_v_count ? ga : gb).push_back(g);
        }
        c->red_n_a = (int) ga.size();
        c->red_n_b_groups = (int) gb.size();
        const int per_cta = kSymV / c->sym_v_count;
        c->red_n_b_ctas = (c->red_n_b_groups + per_cta - 1) / per_cta;
        ga.insert(ga.end(), gb.begin(), gb.end());
        if (c->red_groups) cudaFree(c->red_groups);
        c->red_groups = nullptr;
        CU(dalloc(&c->red_groups, ga.size()));
        CU(cudaMemcpy(c->red_groups, ga.data(), sizeof(int) * ga.size(), cudaMemcpyHostToDevice));
    }
    if (c->sym_trace) cudaFree(c->sym_trace);
    c->sym_trace = nullptr;
    sp.trace = nullptr;
    if (getenv("NPSL_SYM_TRACE")) {  // measurement only: per-CTA timeline of the last k_pair_sym launch
        CU(dalloc(&c->sym_trace, 5 * units.size() + 5));
        sp.trace = c->sym_trace;
        c->sym_trace_units = units;
    }
    return 0;
}

template <int ROWS>
void launch_pair(npsl_ctx *c, bool energy, cudaStream_t s) {
    dim3 grid(c->pair_row_tiles, c->pair_splits);
    if (energy)
        k_pair<ROWS, true><<<grid, kPairThreads, 0, s>>>(c->xyzq, c->pair_partial, c->lj_hit, c->pp);
    else
        k_pair<ROWS, false><<<grid, kPairThreads, 0, s>>>(c->xyzq, c->pair_partial, c->lj_hit, c->pp);
}

int gather_records(npsl_ctx *c, double4 *buf);

// Issues the kernels of one force evaluation (and, with step=true, of one whole MD step) on the
// context's streams. Used both for direct execution and under stream capture.
int issue(npsl_ctx *c, bool step, bool energy, bool store, bool time_pair, bool packed = false) {
    const Packed pk = {packed ? c->d_state_in : nullptr, packed ? c->d_state_out : nullptr};
    const VertexTables vt = {c->degree, c->ring, c->ring_l0, c->bslot_ref};
    const ForceComponents fc = {c->comp[0], c->comp[1], c->comp[2], c->comp[3], c->comp[4], c->vertex_e};
    int n = 0;
    if (!step) c->local_valid = energy && store;
    auto mark = [&](const char *name) {
        if (!c->prof) return;
        if (c->prof_used == c->prof_ev.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            c->prof_ev.push_back(e);
            c->prof_names.push_back(name);
        }
        c->prof_names[c->prof_used] = name;
        cudaEventRecord(c->prof_ev[c->prof_used++], c->s_main);
    };
    mark("begin");
    if (step) {
        k_kick_drift<<<c->n_vertex_blocks, kMeshThreads, 0, c->s_main>>>(c->xyzq, c->vel, c->force, c->scal, c->lj_hit,
                                                                        c->d_peer_xyzq, c->xg, c->mp, pk);
        n++;
        if (packed) {  // the new positions are final: their copy to the host runs beside the force evaluation
            CU(cudaEventRecord(c->ev_copy_fork, c->s_main));
            CU(cudaStreamWaitEvent(c->s_copy, c->ev_copy_fork, 0));
            CU(cudaMemcpyAsync(c->h_state, c->d_state_out, sizeof(double) * 3 * c->nv, cudaMemcpyDeviceToHost, c->s_copy));
            CU(cudaEventRecord(c->ev_copy_join, c->s_copy));
        }
        if (c->mp.n_ions > 0) {
            k_ion_kick_drift<<<(c->mp.n_ions + 127) / 128, 128, 0, c->s_main>>>(c->ion_xyzq, c->ion_vel, c->ion_force, c->scal, c->ip);
            n++;
        }
        mark("k_kick_drift");
        if (c->xg.on & (1 << XK_POS)) {
            k_xchg_wait<<<1, 32, 0, c->s_main>>>(c->xg, XK_POS);
            n++;
        } else if (c->world > 1 && !c->replicated) {
            NC(g_nccl.AllGather(c->xyzq + (size_t) c->rank * c->range, c->xyzq, (size_t) c->range * 4, ncclFloat64_,
                                c->comm, c->s_main));
        }
        if (c->world > 1 && !c->replicated) mark("exchange positions");
        if (c->mp.constraint) {  // SHAKE, src/newmd.cpp:126
            // rows plan: the face sums run over all faces on every rank (identical order, identical multiplier); they
            // need VERTEX::neg_GradVi of every vertex, gathered at the end of the last force evaluation (below)
            k_constraint_sums<0><<<c->n_face_blocks, kMeshThreads, 0, c->s_main>>>(c->xyzq, c->vel, c->ngv, c->face_v,
                                                                                  c->part_con, c->tickets + 2, c->scal, c->mp);
            k_shake_apply<<<(c->nv + kMeshThreads - 1) / kMeshThreads, kMeshThreads, 0, c->s_main>>>(c->xyzq, c->vel, c->ngv,
                                                                                                   c->scal, c->mp);
            n += 2;
            mark("SHAKE");
        }
    } else {
        CU(cudaMemsetAsync(c->lj_hit, 0, sizeof(int), c->s_main));
    }
    // Mesh terms on the low-priority side stream. They are issued after the pair kernel: its CTAs fill every SM
    // (three per SM take all registers), so the mesh CTAs run in the ragged end of the pair launch instead of
    // delaying its start. Uncharged meshes have no pair kernel; there the order does not matter.
    CU(cudaEventRecord(c->ev_fork, c->s_main));
    auto issue_mesh = [&]() -> int {
    CU(cudaStreamWaitEvent(c->s_side, c->ev_fork, 0));
    {
        const int grid = c->n_face_blocks + c->n_edge_blocks;
        unsigned int *tk = c->tickets;
#define MESH(E, S)                                                                                              \
    k_mesh<E, S><<<grid, kMeshThreads, 0, c->s_side>>>(c->xyzq, c->face_v, c->edge_rec, c->l0, c->bslot,         \
                                                       c->part_mesh, tk, c->scal, c->face_dir_area, c->face_vol, \
                                                       c->edge_S, c->edge_E, c->n_face_blocks, c->mp)
        if (energy && store) MESH(true, true);
        else if (energy) MESH(true, false);
        else if (store) MESH(false, true);
        else MESH(false, false);
#undef MESH
        n++;
        if (step)
            k_vertex_mesh<VMODE_STEP><<<c->n_vertex_blocks, kMeshThreads, 0, c->s_side>>>(c->xyzq, c->bslot, vt, fc, c->scal,
                                                                                       c->fmesh, c->ngv, c->mp);
        else
            k_vertex_mesh<VMODE_FORCE><<<c->n_vertex_blocks, kMeshThreads, 0, c->s_side>>>(c->xyzq, c->bslot, vt, fc, c->scal,
                                                                                        c->fmesh, c->ngv, c->mp);
        n++;
    }
    if (c->mp.n_ions > 0) {  // mesh-ion and ion-ion terms, beside the mesh-mesh pair kernel
        const int ion_ctas = (c->nv + 127) / 128;
        k_ion_mesh<<<dim3(ion_ctas, (c->mp.n_ions + 127) / 128), 128, 0, c->s_side>>>(c->xyzq, c->ion_xyzq, c->ion_mesh, c->ion_mesh_lj,
                                                                                        c->ion_part, c->ip);
        k_ion_rows<<<c->mp.n_ions, 256, 0, c->s_side>>>(c->ion_xyzq, c->ion_part, ion_ctas, c->ion_force, c->ion_e, c->ip);
        n += 2;
    }
    CU(cudaEventRecord(c->ev_join, c->s_side));
        return 0;
    };
    bool mesh_issued = false;
    cudaStream_t lj_stream = c->s_main;
    if (!c->all_q_zero) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (time_pair && c->t_used + 2 <= c->t_ev.size()) {
            e0 = c->t_ev[c->t_used];
            e1 = c->t_ev[c->t_used + 1];
            c->t_used += 2;
            CU(cudaEventRecord(e0, c->s_main));
        }
        const bool ordered = !c->use_sym || energy;  // the energy sum always comes from the ordered kernel
        if (c->use_sym) {
#define SYM_LAUNCH(R, T, B, MATH, U, SMEM)                                                                           \
    do {                                                                                                            \
        if (SMEM > 0) CU(cudaFuncSetAttribute(k_pair_sym<R, T, B, MATH, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); \
        k_pair_sym<R, T, B, MATH, U><<<c->sym_n_units, T, SMEM, c->s_main>>>(c->xyzq, c->sym_units, c->sym_rowpart,  \
                                                                             c->sym_colpart, c->lj_hit, c->sp);     \
    } while (0)
            // (math, unroll, rows per thread, threads per CTA, tighter register cap) -> instantiation. The default is
            // SYM_FAST with 6 rows x 64 threads (4 CTAs of 2 warps per SM); the others are kept for A/B measurements
            // (NPSL_SYM_MATH, NPSL_SYM_UNROLL, npsl_set_pair_plan) and as the full-accuracy reference.
            constexpr int kPolySmem = kPolyCount * 48;
            const int key = ((c->sym_math * 10 + c->sym_unroll) * 10 + c->sym_rows) * 10000 + c->sym_threads * 10 + (c->sym_minb >= 8 ? 1 : 0);
#define KEY(M, U, R, T, TIGHT) ((((M) * 10 + (U)) * 10 + (R)) * 10000 + (T) * 10 + (TIGHT))
            switch (key) {
                case KEY(SYM_FAST, 1, 6, 64, 0): SYM_LAUNCH(6, 64, 4, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 1, 6, 32, 0): SYM_LAUNCH(6, 32, 8, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 1, 6, 128, 0): SYM_LAUNCH(6, 128, 2, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 1, 6, 256, 0): SYM_LAUNCH(6, 256, 1, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 1, 4, 128, 0): SYM_LAUNCH(4, 128, 3, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 1, 4, 128, 1): SYM_LAUNCH(4, 128, 4, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 1, 4, 64, 0): SYM_LAUNCH(4, 64, 6, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 1, 5, 128, 0): SYM_LAUNCH(5, 128, 3, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 1, 7, 128, 0): SYM_LAUNCH(7, 128, 2, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 1, 8, 128, 0): SYM_LAUNCH(8, 128, 2, SYM_FAST, 1, 0); break;
                case KEY(SYM_FAST, 2, 4, 128, 0): SYM_LAUNCH(4, 128, 2, SYM_FAST, 2, 0); break;
                case KEY(SYM_POLY, 1, 4, 128, 0): SYM_LAUNCH(4, 128, 3, SYM_POLY, 1, kPolySmem); break;
                case KEY(SYM_EXACT, 1, 4, 128, 0): SYM_LAUNCH(4, 128, 3, SYM_EXACT, 1, 0); break;
                case KEY(SYM_EXACT, 1, 4, 128, 1): SYM_LAUNCH(4, 128, 4, SYM_EXACT, 1, 0); break;
                case KEY(SYM_EXACT, 2, 4, 128, 0): SYM_LAUNCH(4, 128, 2, SYM_EXACT, 2, 0); break;
                case KEY(SYM_EXACT, 1, 6, 128, 0): SYM_LAUNCH(6, 128, 2, SYM_EXACT, 1, 0); break;
                case KEY(SYM_EXACT, 1, 6, 64, 0): SYM_LAUNCH(6, 64, 4, SYM_EXACT, 1, 0); break;
                case KEY(SYM_EXACT, 1, 2, 256, 0): SYM_LAUNCH(2, 256, 3, SYM_EXACT, 1, 0); break;
                case KEY(SYM_EXACT, 1, 2, 128, 0): SYM_LAUNCH(2, 128, 6, SYM_EXACT, 1, 0); break;
                case KEY(SYM_EXACT, 1, 2, 128, 1): SYM_LAUNCH(2, 128, 8, SYM_EXACT, 1, 0); break;
                case KEY(SYM_EXACT, 1, 1, 256, 0): SYM_LAUNCH(1, 256, 4, SYM_EXACT, 1, 0); break;
                case KEY(SYM_EXACT, 1, 1, 128, 0): SYM_LAUNCH(1, 128, 8, SYM_EXACT, 1, 0); break;
                default:
                    return fail(c, NPSL_ERR_ARG, "no k_pair_sym instantiation for math %d, unroll %d, %d rows x %d threads%s",
                                c->sym_math, c->sym_unroll, c->sym_rows, c->sym_threads, c->sym_minb >= 8 ? " (tight)" : "");
            }
#undef KEY
#undef SYM_LAUNCH
            n++;
            if (e1) CU(cudaEventRecord(e1, c->s_main));
            mark("k_pair_sym");
            { int r = issue_mesh(); if (r < 0) return r; mesh_issued = true; }
            if (!ordered && c->world > 1 && !c->p2p) {  // NCCL exchanges: every rank's collision flag, OR-ed (see k_lj_flag)
                NC(g_nccl.AllGather(c->lj_hit, c->lj_all, sizeof(int), 1 /* ncclChar */, c->comm, c->s_main));
                k_lj_or<<<1, 32, 0, c->s_main>>>(c->lj_all, c->world, c->lj_hit);
                n++;
            }
            if (!ordered) {  // the exact LJ pass only needs the collision detector: it runs beside k_pair_reduce
                CU(cudaEventRecord(c->ev_lj_fork, c->s_main));
                CU(cudaStreamWaitEvent(c->s_aux, c->ev_lj_fork, 0));
                lj_stream = c->s_aux;
                if (c->world > 1 && c->p2p) {  // the detector of k_pair_sym only saw this rank's units
                    k_lj_flag<<<1, 32, 0, lj_stream>>>(c->xg, c->lj_hit);
                    n++;
                }
            }
            k_pair_reduce<<<c->red_n_a + c->red_n_b_ctas, dim3(32, 8), 0, c->s_main>>>(c->sym_rowpart, c->sym_colpart, c->sym_fv,
                                                                            c->red_groups, c->red_n_a, c->red_n_b_groups,
                                                                            c->sym_v_first, c->sym_v_count,
                                                                            c->d_peer_fv, c->range, c->xg, c->sp);
            n++;
            if (c->world > 1 && !c->p2p) {
                const size_t per_rank = (size_t) c->sym_v_count * 3 * c->sp.npad;
                NC(g_nccl.AllGather(c->sym_fv + (size_t) c->rank * per_rank, c->sym_fv, per_rank, ncclFloat64_, c->comm,
                                    c->s_main));
            }
            mark(c->world > 1 && !c->p2p ? "k_pair_reduce + all-gather" : "k_pair_reduce");
        }
        if (ordered) {
            switch (c->pair_rows_per_thread) {
                case 1: launch_pair<1>(c, energy, c->s_main); break;
                case 2: launch_pair<2>(c, energy, c->s_main); break;
                default: launch_pair<4>(c, energy, c->s_main); break;
            }
            n++;
            if (e1 && !c->use_sym) CU(cudaEventRecord(e1, c->s_main));
        }
    }
    if (!mesh_issued) { int r = issue_mesh(); if (r < 0) return r; }
    k_lj<<<(c->row_hi - c->row_lo + 127) / 128, 128, 0, lj_stream>>>(c->xyzq, c->lj_out, c->lj_hit,
                                                                     c->all_q_zero ? 1 : 0, c->ljp);
    n++;
    if (lj_stream != c->s_main) {
        CU(cudaEventRecord(c->ev_lj_join, lj_stream));
        CU(cudaStreamWaitEvent(c->s_main, c->ev_lj_join, 0));
    }
    mark("k_pair (ordered) + k_lj");
    CU(cudaStreamWaitEvent(c->s_main, c->ev_join, 0));
    mark("join k_mesh");
    if (c->mp.constraint) {  // rows plan: every rank needs every vertex's new neg_GradVi for RATTLE and the next SHAKE
        int r = gather_records(c, c->ngv);
        if (r) return r;
    }
    if (c->mp.n_ions > 0) {  // before the kernel that advances the thermostat chains: it needs the ions' kinetic energy
        if (step) k_ion_finish<true><<<1, 256, 0, c->s_main>>>(c->ion_vel, c->ion_force, c->ion_e, c->scal, c->ip);
        else k_ion_finish<false><<<1, 256, 0, c->s_main>>>(c->ion_vel, c->ion_force, c->ion_e, c->scal, c->ip);
        n++;
    }
    if (step) {
        k_vertex<VMODE_STEP><<<c->n_vertex_blocks, kMeshThreads, 0, c->s_main>>>(
            c->xyzq, c->vel, c->force, c->fmesh, c->pair_partial, c->sym_fv, c->lj_out, c->lj_hit, fc, c->part_vert,
            c->tickets + 1, c->scal, c->ke_warp, c->n_ke_warps, c->th_mid, c->th_cur, c->d_peer_ke, c->xg, c->mp,
            c->ion_mesh, c->ion_mesh_lj, pk);
        n++;
    } else {
        k_vertex<VMODE_FORCE><<<c->n_vertex_blocks, kMeshThreads, 0, c->s_main>>>(
            c->xyzq, c->vel, c->force, c->fmesh, c->pair_partial, c->sym_fv, c->lj_out, c->lj_hit, fc, c->part_vert,
            c->tickets + 1, c->scal, c->ke_warp, c->n_ke_warps, c->th_mid, c->th_cur, c->d_peer_ke, c->xg, c->mp,
            c->ion_mesh, c->ion_mesh_lj, Packed{nullptr, nullptr});
        n++;
    }
    mark("k_vertex");
    if (step && c->mp.constraint) {  // RATTLE, src/newmd.cpp:152, then the kinetic energy and the thermostat
        { int r = gather_records(c, c->vel); if (r) return r; }  // rows plan: d/dt of the volume needs every velocity
        k_constraint_sums<1><<<c->n_face_blocks, kMeshThreads, 0, c->s_main>>>(c->xyzq, c->vel, c->ngv, c->face_v,
                                                                              c->part_con, c->tickets + 2, c->scal, c->mp);
        k_rattle_apply<false><<<c->n_vertex_blocks, kMeshThreads, 0, c->s_main>>>(c->vel, c->ngv, c->part_vert, c->tickets + 1,
                                                                                  c->scal, c->ke_warp, c->n_ke_warps,
                                                                                  c->th_mid, c->th_cur, c->d_peer_ke, c->xg, c->mp);
        n += 2;
        mark("RATTLE");
    }
    if (c->world > 1 && !c->replicated) {
        const size_t per_rank = (size_t) c->range / 32;
        if (!c->p2p)
            NC(g_nccl.AllGather(c->ke_warp + (size_t) c->rank * per_rank, c->ke_warp, per_rank, ncclFloat64_, c->comm,
                                c->s_main));
        k_ke_post<<<1, kMeshThreads, 0, c->s_main>>>(c->ke_warp, c->n_ke_warps, c->scal, c->th_mid, c->th_cur,
                                                     step ? 1 : 0, c->xg, c->mp);
        mark("exchange KE + k_ke_post");
        n++;
    }
    CU(cudaGetLastError());
    return n;
}

// true when [p, p+bytes) is page-locked: either registered by an earlier call or registered now
bool user_pinned(npsl_ctx *c, void *p, size_t bytes) {
    for (auto &e : c->pinned_user)
        if (e.first == p && e.second >= bytes) return true;
    if (c->pinned_user.size() >= 16) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeHost) return true;  // caller pinned it
    cudaGetLastError();
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    c->pinned_user.emplace_back(p, bytes);
    return true;
}

int ensure_graph(npsl_ctx *c) {
    if (c->step_graph) return 0;
    cudaGraph_t g = nullptr;
    CU(cudaStreamBeginCapture(c->s_main, cudaStreamCaptureModeThreadLocal));
    int n = issue(c, true, false, false, false);
    cudaError_t e = cudaStreamEndCapture(c->s_main, &g);
    if (n < 0) {
        if (g) cudaGraphDestroy(g);
        return n;
    }
    if (e != cudaSuccess) return fail(c, NPSL_ERR_CUDA, "graph capture: %s", cudaGetErrorString(e));
    c->launches_per_step = n;
    e = cudaGraphInstantiateWithFlags(&c->step_graph, g, cudaGraphInstantiateFlagUseNodePriority);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(c, NPSL_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(e));
    return 0;
}

// The packed host-state step is one graph: copy of the caller's block in, the step with k_kick_drift reading and
// k_vertex writing the packed images, the positions copied out beside the force evaluation, velocities + forces +
// kinetic energy copied out at the end. Single-GPU contexts and the replicated plan without constraint / ions (there
// every rank integrates every vertex and the thermostat runs inside k_vertex); otherwise npsl_md_step_packed falls
// back to unpack -> steps -> pack.
bool packed_fused_ok(const npsl_ctx *c) {
    return (c->world == 1 || c->replicated) && !c->mp.constraint && c->mp.n_ions == 0;
}

int ensure_packed_buffers(npsl_ctx *c) {
    if (c->h_state) return 0;
    const size_t n = (size_t) 9 * c->nv + 1;
    CU(cudaMallocHost((void **) &c->h_state, sizeof(double) * n));
    memset(c->h_state, 0, sizeof(double) * n);
    CU(dalloc(&c->d_state_in, n));
    CU(dalloc(&c->d_state_out, n));
    CU(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_copy_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_copy_join, cudaEventDisableTiming));
    return 0;
}

int ensure_packed_graph(npsl_ctx *c) {
    if (c->packed_graph) return 0;
    cudaGraph_t g = nullptr;
    const size_t n3 = (size_t) 3 * c->nv;
    CU(cudaStreamBeginCapture(c->s_main, cudaStreamCaptureModeThreadLocal));
    cudaError_t e0 = cudaSuccess;
    int extra = 0;
    if (c->xg.on & (1 << XK_STATE)) {
        // replicated plan over peer memory: this rank's PCIe link carries its row slice only (the reference's own row
        // decomposition, src/main.cpp:448-455); the slices travel on over NVLink
        const int lo = std::min(c->rank * c->range, c->nv), hi = std::min((c->rank + 1) * c->range, c->nv);
        for (int k = 0; k < 3 && e0 == cudaSuccess && hi > lo; k++)
            e0 = cudaMemcpyAsync(c->d_state_in + k * n3 + (size_t) 3 * lo, c->h_state + k * n3 + (size_t) 3 * lo,
                                 sizeof(double) * 3 * (hi - lo), cudaMemcpyHostToDevice, c->s_main);
        const int total = 9 * std::max(hi - lo, 0);
        k_state_share<<<std::max(1, (total + 255) / 256), 256, 0, c->s_main>>>(c->d_state_in, c->d_peer_state, c->nv, lo,
                                                                              std::max(hi, lo), c->xg);
        extra = 1;
    } else {
        e0 = cudaMemcpyAsync(c->d_state_in, c->h_state, sizeof(double) * 3 * n3, cudaMemcpyHostToDevice, c->s_main);
    }
    int n = e0 == cudaSuccess ? issue(c, true, false, false, false, true) : -1;
    if (n >= 0) n += extra;
    if (n >= 0) {
        cudaStreamWaitEvent(c->s_main, c->ev_copy_join, 0);
        e0 = cudaMemcpyAsync(c->h_state + n3, c->d_state_out + n3, sizeof(double) * (2 * n3 + 1), cudaMemcpyDeviceToHost, c->s_main);
    }
    cudaError_t e = cudaStreamEndCapture(c->s_main, &g);
    if (n < 0 || e0 != cudaSuccess) {
        if (g) cudaGraphDestroy(g);
        return n < 0 && e0 == cudaSuccess ? n : fail(c, NPSL_ERR_CUDA, "packed graph capture: %s", cudaGetErrorString(e0));
    }
    if (e != cudaSuccess) return fail(c, NPSL_ERR_CUDA, "packed graph capture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiateWithFlags(&c->packed_graph, g, cudaGraphInstantiateFlagUseNodePriority);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(c, NPSL_ERR_CUDA, "packed graph instantiate: %s", cudaGetErrorString(e));
    c->packed_launches = n;
    return 0;
}

// A wait on a peer that gave up (xchg.cuh: every rank has to issue the same sequence of force evaluations)
int xchg_status(npsl_ctx *c) {
    if (!c->x_timed_out) return 0;
    int v = 0;
    CU(cudaMemcpy(&v, c->x_timed_out, sizeof(int), cudaMemcpyDeviceToHost));
    if (v)
        return fail(c, NPSL_ERR_STATE, "an exchange with a peer GPU timed out after %llu s: the ranks of a sharded context must "
                    "issue the same sequence of force evaluations (npsl_step, npsl_force*, npsl_energy, ...); results are invalid",
                    (unsigned long long) (kXchgTimeoutNs / 1000000000ull));
    return 0;
}

int collect_timing(npsl_ctx *c) {
    if (!c->t_used) return 0;
    CU(cudaStreamSynchronize(c->s_main));
    for (size_t k = 0; k + 1 < c->t_used; k += 2) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->t_ev[k], c->t_ev[k + 1]));
        c->t_sum_ms += ms;
        c->t_count++;
    }
    c->t_used = 0;
    return 0;
}

// in-place all-gather of a per-vertex record array laid out as world*range records
int gather_records(npsl_ctx *c, double4 *buf) {
    if (c->world == 1 || c->replicated) return 0;  // replicated integrator: every rank holds every vertex
    NC(g_nccl.AllGather(buf + (size_t) c->rank * c->range, buf, (size_t) c->range * 4, ncclFloat64_, c->comm,
                        c->s_main));
    return 0;
}

// records on the device -> packed [nv][3] host array, through the pinned staging buffer `slot`
int download_records(npsl_ctx *c, double4 *dev, double *out, bool gather, int slot = 0) {
    if (!out) return 0;
    if (gather) {
        int r = gather_records(c, dev);
        if (r) return r;
    }
    const int nb = (c->nv + 255) / 256;
    k_gather3<<<nb, 256, 0, c->s_main>>>(dev, c->d_pack[slot], c->nv);
    c->launches++;
    CU(cudaMemcpyAsync(c->h_pin[slot], c->d_pack[slot], sizeof(double) * 3 * c->nv, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    memcpy(out, c->h_pin[slot], sizeof(double) * 3 * c->nv);
    return 0;
}

// packed [nv][3] host array -> xyz of the device records (w kept or zeroed)
int upload_records(npsl_ctx *c, double4 *dev, const double *in, bool keep_w, int slot = 0) {
    if (!in) return 0;
    CU(cudaStreamSynchronize(c->s_main));  // the staging buffer may still be in flight
    memcpy(c->h_pin[slot], in, sizeof(double) * 3 * c->nv);
    CU(cudaMemcpyAsync(c->d_pack[slot], c->h_pin[slot], sizeof(double) * 3 * c->nv, cudaMemcpyHostToDevice, c->s_main));
    const int nb = (c->nv + 255) / 256;
    k_scatter3<<<nb, 256, 0, c->s_main>>>(dev, c->d_pack[slot], c->nv, keep_w ? 1 : 0);
    c->launches++;
    return 0;
}

// Rows plan over peer memory: the ranks only meet inside a step, through the exchanges. A host-state call rewrites every
// position of this rank's copy (and later reads them all back), while a rank that is a call ahead would already push
// the next step's positions into that copy; one tiny NCCL collective after the rewrite and one after the read-back
// keep the calls of the ranks in lock-step.
int rows_barrier(npsl_ctx *c) {
    if (c->world == 1 || c->replicated || !c->p2p) return 0;
    NC(g_nccl.AllGather(c->scal, c->scal_all, SC_COUNT, ncclFloat64_, c->comm, c->s_main));
    return 0;
}

void drop_graph(npsl_ctx *c) {
    if (c->step_graph) {
        cudaGraphExecDestroy(c->step_graph);
        c->step_graph = nullptr;
    }
    if (c->packed_graph) {
        cudaGraphExecDestroy(c->packed_graph);
        c->packed_graph = nullptr;
    }
}

// Maps every rank's exchange buffers (positions, virtual-rank force sums, kinetic-energy partials, arrival
// flags) into this process with CUDA IPC. The handles travel through one NCCL all-gather. Any failure leaves
// the context on the NCCL exchanges.
int setup_p2p(npsl_ctx *c) {
    c->xg.world = c->world; c->xg.rank = c->rank; c->xg.on = 0; c->xg.fv_bcast = 0; c->xg.cta_fence_sys = 0;
    c->xg.peer_flags = nullptr; c->xg.state = nullptr; c->xg.tickets = nullptr; c->xg.timed_out = nullptr;
    if (c->world == 1) return 0;
    CU(dalloc(&c->lj_all, c->world));
    const char *env = getenv("NPSL_EXCHANGE");
    if (env && strcmp(env, "nccl") == 0) return 0;
    if (c->world > kMaxRanks || !c->sym_fv) return 0;
    CU(dalloc(&c->xflags, XK_KINDS * kMaxRanks));
    CU(dalloc(&c->xstate, 2 * XK_KINDS));
    CU(dalloc(&c->xtickets, XK_KINDS));
    CU(dalloc(&c->x_timed_out, 1));
    {  // the packed image of npsl_md_step_packed is shared slice by slice (k_state_share): it has to exist before the mapping
        int r = ensure_packed_buffers(c);
        if (r) return r;
    }
    struct Handles { cudaIpcMemHandle_t h[5]; int ok; int pad[15]; };
    Handles mine;
    memset(&mine, 0, sizeof mine);
    void *bufs[5] = {c->xyzq, c->sym_fv, c->ke_warp, c->xflags, c->d_state_in};
    mine.ok = 1;
    for (int k = 0; k < 5; k++)
        if (cudaIpcGetMemHandle(&mine.h[k], bufs[k]) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
    Handles *d_all = nullptr;
    CU(dalloc(&d_all, c->world));
    CU(cudaMemcpy(d_all + c->rank, &mine, sizeof mine, cudaMemcpyHostToDevice));
    NC(g_nccl.AllGather(d_all + c->rank, d_all, sizeof(Handles), 1 /* ncclChar */, c->comm, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    std::vector<Handles> all(c->world);
    CU(cudaMemcpy(all.data(), d_all, sizeof(Handles) * c->world, cudaMemcpyDeviceToHost));
    cudaFree(d_all);
    bool ok = true;
    for (int r = 0; r < c->world; r++) ok = ok && all[r].ok;
    for (int r = 0; r < c->world && ok; r++)
        for (int k = 0; k < 5; k++) {
            if (r == c->rank) { c->peer_map[k][r] = bufs[k]; continue; }
            if (cudaIpcOpenMemHandle(&c->peer_map[k][r], all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ok = false;
                cudaGetLastError();
                c->peer_map[k][r] = nullptr;
                break;
            }
        }
    // every rank must take the same decision: agree through a second tiny all-gather
    int *d_ok = nullptr;
    CU(dalloc(&d_ok, c->world));
    int my_ok = ok ? 1 : 0;
    CU(cudaMemcpy(d_ok + c->rank, &my_ok, sizeof(int), cudaMemcpyHostToDevice));
    NC(g_nccl.AllGather(d_ok + c->rank, d_ok, sizeof(int), 1, c->comm, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    std::vector<int> oks(c->world);
    CU(cudaMemcpy(oks.data(), d_ok, sizeof(int) * c->world, cudaMemcpyDeviceToHost));
    cudaFree(d_ok);
    for (int r = 0; r < c->world; r++) ok = ok && oks[r];
    if (!ok) return 0;
    CU(dalloc(&c->d_peer_xyzq, c->world)); CU(dalloc(&c->d_peer_fv, c->world));
    CU(dalloc(&c->d_peer_ke, c->world)); CU(dalloc(&c->d_peer_flags, c->world));
    CU(cudaMemcpy(c->d_peer_xyzq, c->peer_map[0], sizeof(void *) * c->world, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_peer_fv, c->peer_map[1], sizeof(void *) * c->world, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_peer_ke, c->peer_map[2], sizeof(void *) * c->world, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_peer_flags, c->peer_map[3], sizeof(void *) * c->world, cudaMemcpyHostToDevice));
    CU(dalloc(&c->d_peer_state, c->world));
    CU(cudaMemcpy(c->d_peer_state, c->peer_map[4], sizeof(void *) * c->world, cudaMemcpyHostToDevice));
    c->xg.peer_flags = c->d_peer_flags; c->xg.state = c->xstate; c->xg.tickets = c->xtickets;
    c->xg.timed_out = c->x_timed_out;
    c->p2p = true;
    c->xg.on = (1 << XK_POS) | (1 << XK_FV) | (1 << XK_KE);
    return 0;
}

// Chooses between the two sharding plans of a multi-GPU context and sets the owned ranges, launch geometry and
// exchange mask accordingly. Every rank takes the same decision (it depends on the vertex count, the charges,
// the environment and on whether peer mapping succeeded, which the ranks agreed on in setup_p2p).
//   rows        vertices are owned in slices (src/main.cpp:448-455); positions, pair-force sums and kinetic-energy
//               partials are exchanged every step. Less redundant O(N) work, three exchange latencies per step.
//   replicated  only the pair term is split (the reference's own scheme: it all-gathers forces and lets every rank
//               integrate, src/newforces.cpp:268-271); one exchange per step. Wins while the O(N) kernels are
//               latency-bound, i.e. for meshes up to about 1e5 vertices.
constexpr int kReplicateUpTo = 100000;
int apply_shard_plan(npsl_ctx *c) {
    const bool can = c->world > 1 && c->p2p && c->use_sym && !c->all_q_zero;
    const bool want = c->shard_pref == 2 || (c->shard_pref == 0 && c->nv <= kReplicateUpTo);
    c->replicated = can && want;
    if (c->replicated) {
        c->row_lo = 0;
        c->row_hi = c->nv;
    } else {
        c->row_lo = std::min(c->rank * c->range, c->nv);
        c->row_hi = std::min((c->rank + 1) * c->range, c->nv);
    }
    c->n_vertex_blocks = (c->row_hi - c->row_lo + kMeshThreads - 1) / kMeshThreads;
    c->pp.row_lo = c->ljp.row_lo = c->mp.row_lo = c->row_lo;
    c->pp.row_hi = c->ljp.row_hi = c->mp.row_hi = c->row_hi;
    c->mp.n_ranks = c->replicated ? 1 : c->world;
    c->xg.on = !c->p2p ? 0 : (c->replicated ? (1 << XK_FV) | (1 << XK_STATE) : (1 << XK_POS) | (1 << XK_FV) | (1 << XK_KE));
    c->xg.fv_bcast = c->replicated ? 1 : 0;
    c->xg.cta_fence_sys = 0;
    if (const char *e = getenv("NPSL_CTA_FENCE_SYS")) c->xg.cta_fence_sys = atoi(e) ? 1 : 0;  // measurement knob
    CU(cudaStreamSynchronize(c->s_main));
    plan_pair(c, c->n_sm);
    if (c->pair_partial) cudaFree(c->pair_partial);
    c->pair_partial = nullptr;
    CU(dalloc(&c->pair_partial, (size_t) c->pair_splits * c->pair_row_stride));
    apply_pair_plan(c);
    drop_graph(c);
    c->forces_valid = false;
    return 0;
}

int create_impl(const npsl_mesh_desc *mesh, const npsl_params *prm, int rank, int world, const void *nccl_id,
                npsl_ctx **out) {
    npsl_ctx *c = nullptr;
    if (!out) return fail(c, NPSL_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!mesh || !prm || !mesh->edge_v || !mesh->face_v || !mesh->l0) return fail(c, NPSL_ERR_ARG, "NULL argument");
    if (mesh->n_vertices < 4 || mesh->n_edges < 6 || mesh->n_faces < 4) return fail(c, NPSL_ERR_ARG, "mesh too small");
    if (world < 1 || rank < 0 || rank >= world) return fail(c, NPSL_ERR_ARG, "bad rank/world");
    if (prm->chain < 1 || prm->chain > NPSL_MAX_CHAIN) return fail(c, NPSL_ERR_ARG, "chain length out of range");
    if (!(prm->inv_kappa_out > 0) || !(prm->em > 0)) return fail(c, NPSL_ERR_ARG, "inv_kappa_out and em must be positive");
    if (world > 1 && !nccl_id) return fail(c, NPSL_ERR_ARG, "sharded context needs an NCCL unique id");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(c, NPSL_ERR_CUDA, "no CUDA device (%s); this library has no CPU path",
                    ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
    Tables tb;
    std::string why;
    if (build_tables(mesh, tb, why)) return fail(c, NPSL_ERR_ARG, "unsupported mesh: %s", why.c_str());

    c = new npsl_ctx();
    npsl_ctx *guard = c;
    auto bail = [&](int code) { std::string e = c->err; npsl_destroy(guard); g_error = e; return code; };
#define CUB(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            fail(c, NPSL_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                      \
            return bail(NPSL_ERR_CUDA);                                                           \
        }                                                                                         \
    } while (0)
    CUB(cudaGetDevice(&c->device));
    cudaDeviceProp prop;
    CUB(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major < 10) {
        fail(c, NPSL_ERR_CUDA, "device %s is sm_%d%d; this build targets sm_100a only", prop.name, prop.major, prop.minor);
        return bail(NPSL_ERR_CUDA);
    }
    c->nv = mesh->n_vertices; c->ne = mesh->n_edges; c->nf = mesh->n_faces;
    c->rank = rank; c->world = world;
    c->range = (c->nv + world - 1) / world;  // src/main.cpp:448
    // rounded up to a whole warp of vertices: keeps the per-warp kinetic-energy partials (and so the
    // trajectory) bit-identical to the single-GPU run; a rank may then own a few rows more or less
    c->range = (c->range + 31) / 32 * 32;
    c->row_lo = std::min(rank * c->range, c->nv);
    c->row_hi = std::min((rank + 1) * c->range, c->nv);
    if (c->row_hi <= c->row_lo) {
        fail(c, NPSL_ERR_ARG, "rank %d of %d owns no rows of %d vertices", rank, world, c->nv);
        return bail(NPSL_ERR_ARG);
    }
    c->prm = *prm;
    const size_t nrec = (size_t) c->range * world;
    CUB(dalloc(&c->xyzq, nrec)); CUB(dalloc(&c->vel, nrec)); CUB(dalloc(&c->force, nrec)); CUB(dalloc(&c->fmesh, nrec)); CUB(dalloc(&c->ngv, nrec));
    for (int k = 0; k < 5; k++) CUB(dalloc(&c->comp[k], nrec));
    CUB(dalloc(&c->face_v, c->nf)); CUB(dalloc(&c->edge_rec, c->ne)); CUB(dalloc(&c->l0, c->ne));
    CUB(dalloc(&c->degree, c->nv)); CUB(dalloc(&c->ring, (size_t) kMaxDeg * c->nv));
    CUB(dalloc(&c->ring_l0, (size_t) kMaxDeg * c->nv)); CUB(dalloc(&c->bslot_ref, (size_t) 2 * kMaxDeg * c->nv));
    CUB(dalloc(&c->bslot, (size_t) 4 * c->ne));
    CUB(dalloc(&c->lj_out, nrec)); CUB(dalloc(&c->lj_hit, 1));
    CUB(dalloc(&c->scal, SC_COUNT)); CUB(dalloc(&c->scal_all, (size_t) SC_COUNT * world));
    c->n_ke_warps = (int) (nrec / 32);
    CUB(dalloc(&c->ke_warp, c->n_ke_warps));
    CUB(dalloc(&c->th_cur, 1)); CUB(dalloc(&c->th_mid, 1));
    CUB(dalloc(&c->tickets, 4));
    CUB(dalloc(&c->face_dir_area, c->nf)); CUB(dalloc(&c->face_vol, c->nf)); CUB(dalloc(&c->edge_S, c->ne));
    CUB(dalloc(&c->edge_E, c->ne)); CUB(dalloc(&c->vertex_e, nrec)); CUB(dalloc(&c->face_e, c->nf));
    CUB(dalloc(&c->face_local, c->nf)); CUB(dalloc(&c->vertex_geo, c->nv));
    CUB(cudaMemcpy(c->face_e, tb.face_e.data(), sizeof(int4) * c->nf, cudaMemcpyHostToDevice));
    CUB(dalloc(&c->peak_out, 1));
    CUB(cudaMemcpy(c->face_v, tb.face_v.data(), sizeof(int4) * c->nf, cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(c->edge_rec, tb.edge_rec.data(), sizeof(int4) * c->ne, cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(c->l0, mesh->l0, sizeof(double) * c->ne, cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(c->degree, tb.degree.data(), sizeof(int) * c->nv, cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(c->ring, tb.ring.data(), sizeof(int) * tb.ring.size(), cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(c->ring_l0, tb.ring_l0.data(), sizeof(double) * tb.ring_l0.size(), cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(c->bslot_ref, tb.bslot_ref.data(), sizeof(int) * tb.bslot_ref.size(), cudaMemcpyHostToDevice));

    for (int k = 0; k < 3; k++) {
        CUB(cudaMallocHost((void **) &c->h_pin[k], sizeof(double) * 3 * c->nv));
        CUB(dalloc(&c->d_pack[k], (size_t) 3 * c->nv));
    }
    c->n_sm = prop.multiProcessorCount;
    plan_pair(c, c->n_sm);
    CUB(dalloc(&c->pair_partial, (size_t) c->pair_splits * c->pair_row_stride));
    c->n_face_blocks = (c->nf + kMeshThreads - 1) / kMeshThreads;
    c->n_edge_blocks = (c->ne + kMeshThreads - 1) / kMeshThreads;
    c->n_vertex_blocks = (c->row_hi - c->row_lo + kMeshThreads - 1) / kMeshThreads;
    CUB(dalloc(&c->part_mesh, (size_t) 4 * (c->n_face_blocks + c->n_edge_blocks)));
    CUB(dalloc(&c->part_vert, (size_t) 3 * ((nrec + kMeshThreads - 1) / kMeshThreads)));
    CUB(dalloc(&c->part_con, (size_t) 8 * c->n_face_blocks));

    const double lam = prm->inv_kappa_out;
    c->pp.c2s = -1.4426950408889634074 / lam * (double) kExpTab;
    c->pp.inv_lambda = 1.0 / lam;
    c->pp.n = c->nv; c->pp.row_lo = c->row_lo; c->pp.row_hi = c->row_hi;
    const double d2 = prm->lj_length_mesh_mesh * prm->lj_length_mesh_mesh;
    const double cut2 = 1.25992104989487316476 * d2;  // dcut2, src/utility.h:37
    c->pp.lj_cut_hi = high_word(cut2) + 1;             // conservative: exact test happens in k_lj
    c->ljp.d2 = d2; c->ljp.cut2 = cut2; c->ljp.elj = prm->elj;
    c->ljp.n = c->nv; c->ljp.row_lo = c->row_lo; c->ljp.row_hi = c->row_hi;
    MeshParams &mp = c->mp;
    mp.nv = c->nv; mp.ne = c->ne; mp.nf = c->nf; mp.row_lo = c->row_lo; mp.row_hi = c->row_hi;
    mp.n_ranks = world; mp.chain = prm->chain; mp.buckling = prm->buckling ? 1 : 0;
    mp.use_pair = 0;
    mp.constraint = 0;
    mp.n_ions = 0;
    memset(&c->ip, 0, sizeof c->ip);
    apply_pair_plan(c);
    mp.dt = prm->dt; mp.pair_coef = prm->scalefactor / prm->em;
    mp.bkappa = prm->bkappa; mp.sconstant = prm->sconstant; mp.sigma_a = prm->sigma_a; mp.sigma_v = prm->sigma_v;
    mp.ref_volume = prm->ref_volume; mp.avg_edge = prm->avg_edge_length;
    Thermo th;
    memset(&th, 0, sizeof th);
    fill_thermo(th, prm->mesh_bath, prm->ions_bath);
    CUB(cudaMemcpy(c->th_cur, &th, sizeof th, cudaMemcpyHostToDevice));
    CUB(cudaMemcpy(c->th_mid, &th, sizeof th, cudaMemcpyHostToDevice));

    if (plan_sym(c)) return bail(NPSL_ERR_CUDA);
    if (!c->sym_fv) CUB(dalloc(&c->sym_fv, (size_t) 2 * kSymV * 3 * c->sp.npad));
    {
        int prio_least = 0, prio_greatest = 0;
        CUB(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
        CUB(cudaStreamCreateWithPriority(&c->s_main, cudaStreamNonBlocking, prio_greatest));
        CUB(cudaStreamCreateWithPriority(&c->s_side, cudaStreamNonBlocking, prio_least));
        CUB(cudaStreamCreateWithPriority(&c->s_aux, cudaStreamNonBlocking, prio_least));
    }
    CUB(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_lj_fork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_lj_join, cudaEventDisableTiming));
#undef CUB
    if (world > 1) {
        std::string e;
        if (!load_nccl(e)) { fail(c, NPSL_ERR_NCCL, "%s", e.c_str()); return bail(NPSL_ERR_NCCL); }
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof id);
        int r = g_nccl.CommInitRank(&c->comm, world, id, rank);
        if (r != 0) { fail(c, NPSL_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); return bail(NPSL_ERR_NCCL); }
    }
    {
        int r = setup_p2p(c);
        if (r) return bail(r);
        const char *env = getenv("NPSL_SHARD");
        c->shard_pref = env && strcmp(env, "rows") == 0 ? 1 : (env && strcmp(env, "replicated") == 0 ? 2 : 0);
        if ((r = apply_shard_plan(c))) return bail(r);
    }
    *out = c;
    return NPSL_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

int npsl_abi_version(void) { return NPSL_ABI_VERSION; }

const char *npsl_last_error(const npsl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_error.c_str(); }

int npsl_create(const npsl_mesh_desc *mesh, const npsl_params *params, npsl_ctx **out) {
    return create_impl(mesh, params, 0, 1, nullptr, out);
}

int npsl_create_sharded(const npsl_mesh_desc *mesh, const npsl_params *params, int rank, int world,
                        const void *nccl_unique_id, npsl_ctx **out) {
    return create_impl(mesh, params, rank, world, nccl_unique_id, out);
}

int npsl_nccl_unique_id(void *out_id) {
    npsl_ctx *c = nullptr;
    if (!out_id) return fail(c, NPSL_ERR_ARG, "out_id is NULL");
    std::string e;
    if (!load_nccl(e)) return fail(c, NPSL_ERR_NCCL, "%s", e.c_str());
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(out_id, &id, sizeof id);
    return NPSL_OK;
}

void npsl_destroy(npsl_ctx *c) {
    if (!c) return;
    if (c->s_main) cudaStreamSynchronize(c->s_main);
    if (c->sym_trace && getenv("NPSL_SYM_TRACE")) {  // "block I chunk sm start_ns end_ns ready_ns pairs_done_ns" of the last launch
        std::vector<unsigned long long> tr(5 * c->sym_trace_units.size());
        if (cudaMemcpy(tr.data(), c->sym_trace, tr.size() * 8, cudaMemcpyDeviceToHost) == cudaSuccess) {
            if (FILE *f = fopen(getenv("NPSL_SYM_TRACE"), "w")) {
                for (size_t k = 0; k < c->sym_trace_units.size(); k++)
                    fprintf(f, "%zu %d %d %llu %llu %llu %llu %llu\n", k, c->sym_trace_units[k].x, c->sym_trace_units[k].y, tr[5 * k],
                            tr[5 * k + 1], tr[5 * k + 2], tr[5 * k + 3], tr[5 * k + 4]);
                fclose(f);
            }
        }
        cudaFree(c->sym_trace);
    }
    if (c->s_side) cudaStreamSynchronize(c->s_side);
    if (c->s_aux) cudaStreamSynchronize(c->s_aux);
    if (c->step_graph) cudaGraphExecDestroy(c->step_graph);
    if (c->packed_graph) cudaGraphExecDestroy(c->packed_graph);
    if (c->s_copy) { cudaStreamSynchronize(c->s_copy); cudaStreamDestroy(c->s_copy); }
    if (c->ev_copy_fork) cudaEventDestroy(c->ev_copy_fork);
    if (c->ev_copy_join) cudaEventDestroy(c->ev_copy_join);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->d_state_in) cudaFree(c->d_state_in);
    if (c->d_state_out) cudaFree(c->d_state_out);
    if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
    for (auto &e : c->pinned_user) cudaHostUnregister(e.first);
    for (int k = 0; k < 5; k++)
        for (int r = 0; r < kMaxRanks; r++)
            if (c->peer_map[k][r] && r != c->rank) cudaIpcCloseMemHandle(c->peer_map[k][r]);
    for (cudaEvent_t e : c->t_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->s_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    for (int k = 0; k < 3; k++) {
        if (c->h_pin[k]) cudaFreeHost(c->h_pin[k]);
        if (c->d_pack[k]) cudaFree(c->d_pack[k]);
    }
    if (c->flush_buf) cudaFree(c->flush_buf);
    void *ptrs[] = {c->xyzq, c->vel, c->force, c->fmesh, c->ngv, c->part_con, c->comp[0], c->comp[1], c->comp[2], c->comp[3], c->comp[4], c->face_v,
                    c->edge_rec, c->l0, c->degree, c->ring, c->bslot_ref, c->ring_l0, c->bslot, c->pair_partial,
                    c->lj_out, c->lj_hit, c->scal, c->scal_all, c->ke_warp, c->th_cur, c->th_mid, c->part_mesh,
                    c->part_vert, c->tickets, c->ion_xyzq, c->ion_vel, c->ion_force, c->ion_mesh, c->ion_e, c->ion_mesh_lj, c->ion_part, c->edge_E, c->vertex_e, c->face_e, c->face_local, c->vertex_geo, c->face_dir_area, c->face_vol, c->edge_S, c->peak_out, c->sym_units, c->red_groups, c->pair_ptab, c->pair_etab,
                    c->sym_rowpart, c->sym_colpart, c->sym_fv, c->xflags, c->xstate, c->xtickets, c->x_timed_out, c->lj_all,
                    c->d_peer_xyzq, c->d_peer_fv, c->d_peer_ke, c->d_peer_flags, c->d_peer_state};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    if (c->ev_lj_fork) cudaEventDestroy(c->ev_lj_fork);
    if (c->ev_lj_join) cudaEventDestroy(c->ev_lj_join);
    if (c->s_main) cudaStreamDestroy(c->s_main);
    if (c->s_side) cudaStreamDestroy(c->s_side);
    if (c->s_aux) cudaStreamDestroy(c->s_aux);
    delete c;
}

// ---- state ----------------------------------------------------------------------------------------

int npsl_upload_state(npsl_ctx *c, const double *pos, const double *vel, const double *q) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    const int n = c->nv;
    int r;
    if ((r = upload_records(c, c->xyzq, pos, true, 0))) return r;
    if (pos) c->forces_valid = c->local_valid = false;
    if (pos && c->use_sym && c->sym_math == SYM_FAST) {
        // The reduced-accuracy variant of k_pair_sym evaluates 2^(-r / (lambda ln 2)) without an underflow guard: fine
        // up to r = 700 Debye lengths. The reference's meshes have radius 1 in reduced units and lambda ~ 0.3, i.e.
        // r / lambda < 10; a state whose extent comes anywhere near the limit is handed to the exact variant instead.
        double lo[3] = {pos[0], pos[1], pos[2]}, hi[3] = {pos[0], pos[1], pos[2]};
        for (int i = 1; i < n; i++)
            for (int k = 0; k < 3; k++) {
                lo[k] = std::min(lo[k], pos[3 * i + k]);
                hi[k] = std::max(hi[k], pos[3 * i + k]);
            }
        const double extent = std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
        if (!(extent < 100.0 * c->prm.inv_kappa_out)) {
            CU(cudaStreamSynchronize(c->s_main));
            c->sym_math = SYM_EXACT;
            c->sym_shape_set = false;
            if ((r = plan_sym(c))) return r;
            if (c->world > 1 && (r = apply_shard_plan(c))) return r;
            drop_graph(c);
        }
    }
    if ((r = upload_records(c, c->vel, vel, false, 1))) return r;
    if (q) {
        CU(cudaStreamSynchronize(c->s_main));
        memcpy(c->h_pin[0], q, sizeof(double) * n);
        CU(cudaMemcpyAsync(c->d_pack[0], c->h_pin[0], sizeof(double) * n, cudaMemcpyHostToDevice, c->s_main));
        k_scatter_w<<<(n + 255) / 256, 256, 0, c->s_main>>>(c->xyzq, c->d_pack[0], n);
        c->launches++;
        c->all_q_zero = true;
        c->local_valid = false;
        for (int i = 0; i < n; i++)
            if (q[i] != 0.0) { c->all_q_zero = false; break; }
        const int use = c->all_q_zero ? 0 : 1;
        if (use != c->mp.use_pair) {
            c->mp.use_pair = use;
            drop_graph(c);
            if (c->world > 1) {  // uncharged meshes have no pair term to split: rows plan
                int r2 = apply_shard_plan(c);
                if (r2) return r2;
            }
        }
        c->forces_valid = false;
    }
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    return NPSL_OK;
}

int npsl_download_state(npsl_ctx *c, double *pos, double *vel, double *force) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    int r;
    if ((r = download_records(c, c->xyzq, pos, false, 0))) return r;
    if ((r = download_records(c, c->vel, vel, true, 1))) return r;
    if ((r = download_records(c, c->force, force, true, 0))) return r;
    if ((r = xchg_status(c))) return r;
    return collect_timing(c);
}

int npsl_download_force_components(npsl_ctx *c, double *pair, double *bending, double *stretching, double *tension,
                                   double *volume_tension) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    if (!c->forces_valid) return fail(c, NPSL_ERR_STATE, "no force evaluation at the current positions");
    double *outs[5] = {pair, bending, stretching, tension, volume_tension};
    for (int k = 0; k < 5; k++) {
        int r = download_records(c, c->comp[k], outs[k], true);
        if (r) return r;
    }
    return NPSL_OK;
}

int npsl_download_mesh_fields(npsl_ctx *c, double *face_area, double *face_volume, double *face_dir, double *edge_S) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    if (face_area || face_dir) {
        std::vector<double4> tmp(c->nf);
        CU(cudaMemcpyAsync(tmp.data(), c->face_dir_area, sizeof(double4) * c->nf, cudaMemcpyDeviceToHost, c->s_main));
        CU(cudaStreamSynchronize(c->s_main));
        for (int f = 0; f < c->nf; f++) {
            if (face_area) face_area[f] = tmp[f].w;
            if (face_dir) { face_dir[3 * f] = tmp[f].x; face_dir[3 * f + 1] = tmp[f].y; face_dir[3 * f + 2] = tmp[f].z; }
        }
    }
    if (face_volume) CU(cudaMemcpyAsync(face_volume, c->face_vol, sizeof(double) * c->nf, cudaMemcpyDeviceToHost, c->s_main));
    if (edge_S) CU(cudaMemcpyAsync(edge_S, c->edge_S, sizeof(double) * c->ne, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    return NPSL_OK;
}

int npsl_get_thermostat(npsl_ctx *c, npsl_bath_link *mesh_bath, npsl_bath_link *ions_bath) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    Thermo th;
    CU(cudaMemcpyAsync(&th, c->th_cur, sizeof th, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    for (int j = 0; j < c->prm.chain; j++) {
        if (mesh_bath) mesh_bath[j] = {th.mesh[j].Q, th.mesh[j].T, th.mesh[j].dof, th.mesh[j].xi, th.mesh[j].eta};
        if (ions_bath) ions_bath[j] = {th.ions[j].Q, th.ions[j].T, th.ions[j].dof, th.ions[j].xi, th.ions[j].eta};
    }
    return NPSL_OK;
}

int npsl_set_thermostat(npsl_ctx *c, const npsl_bath_link *mesh_bath, const npsl_bath_link *ions_bath) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    Thermo th;
    CU(cudaMemcpyAsync(&th, c->th_cur, sizeof th, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    for (int j = 0; j < c->prm.chain; j++) {
        if (mesh_bath) th.mesh[j] = {mesh_bath[j].Q, mesh_bath[j].T, mesh_bath[j].dof, mesh_bath[j].xi, mesh_bath[j].eta};
        if (ions_bath) th.ions[j] = {ions_bath[j].Q, ions_bath[j].T, ions_bath[j].dof, ions_bath[j].xi, ions_bath[j].eta};
    }
    CU(cudaMemcpyAsync(c->th_cur, &th, sizeof th, cudaMemcpyHostToDevice, c->s_main));
    k_thermo_prepare<<<1, 32, 0, c->s_main>>>(c->th_cur, c->th_mid, c->scal, c->mp);  // reverse half-chain of the next step
    c->launches++;
    CU(cudaStreamSynchronize(c->s_main));
    return NPSL_OK;
}

// ---- the path -------------------------------------------------------------------------------------

int npsl_dressup(npsl_ctx *c, double *total_area, double *total_volume) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    k_mesh<false, true><<<c->n_face_blocks + c->n_edge_blocks, kMeshThreads, 0, c->s_main>>>(
        c->xyzq, c->face_v, c->edge_rec, c->l0, c->bslot, c->part_mesh, c->tickets, c->scal, c->face_dir_area,
        c->face_vol, c->edge_S, c->edge_E, c->n_face_blocks, c->mp);
    c->launches++;
    double s[2];
    CU(cudaMemcpyAsync(s, c->scal + SC_AREA, sizeof s, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    if (total_area) *total_area = s[0];
    if (total_volume) *total_volume = s[1];
    return NPSL_OK;
}

static int force_pass(npsl_ctx *c) {
    int n = issue(c, false, true, true, false);
    if (n < 0) return n;
    c->launches += n;
    c->forces_valid = true;
    return NPSL_OK;
}

int npsl_force_init(npsl_ctx *c) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    return force_pass(c);
}

int npsl_force(npsl_ctx *c) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    return force_pass(c);
}

int npsl_step(npsl_ctx *c, int n_steps) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    if (n_steps < 0) return fail(c, NPSL_ERR_ARG, "n_steps < 0");
    if (!c->forces_valid) return fail(c, NPSL_ERR_STATE, "npsl_step before npsl_force_init");
    if (n_steps > 0) c->local_valid = false;
    if (c->timing) {
        for (int k = 0; k < n_steps; k++) {
            if (c->t_used + 2 > c->t_ev.size()) {
                int r = collect_timing(c);
                if (r) return r;
            }
            int n = issue(c, true, false, false, true);
            if (n < 0) return n;
            c->launches += n;
        }
        return NPSL_OK;
    }
    int r = ensure_graph(c);
    if (r) return r;
    for (int k = 0; k < n_steps; k++) CU(cudaGraphLaunch(c->step_graph, c->s_main));
    c->launches += (int64_t) c->launches_per_step * n_steps;
    return NPSL_OK;
}

int npsl_energy(npsl_ctx *c, npsl_energies *out) {
    if (!c || !out) return fail(c, NPSL_ERR_ARG, "NULL argument");
    if (!c->forces_valid) return fail(c, NPSL_ERR_STATE, "npsl_energy before npsl_force_init");
    // INTERFACE::compute_energy re-walks the pair loop at the current positions and reuses the
    // cached EDGE::itsS / face areas (src/interface.cpp:990-1065); positions have not moved since
    // the last force evaluation, so one energy-enabled force pass yields identical values.
    int r = force_pass(c);
    if (r) return r;
    std::vector<double> all((size_t) SC_COUNT * c->world);
    const int n_parts = c->replicated ? 1 : c->world;  // replicated integrator: this rank walked all rows itself
    if (n_parts > 1) {
        NC(g_nccl.AllGather(c->scal, c->scal_all, SC_COUNT, ncclFloat64_, c->comm, c->s_main));
        CU(cudaMemcpyAsync(all.data(), c->scal_all, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, c->s_main));
    } else {
        CU(cudaMemcpyAsync(all.data(), c->scal, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, c->s_main));
    }
    Thermo th;
    CU(cudaMemcpyAsync(&th, c->th_cur, sizeof th, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    double ke = 0, es = 0, lj = 0;
    for (int k = 0; k < n_parts; k++) {  // fixed rank order
        es += all[(size_t) k * SC_COUNT + SC_ES];
        lj += all[(size_t) k * SC_COUNT + SC_LJ];
    }
    const double *s = all.data() + (size_t) (n_parts > 1 ? c->rank : 0) * SC_COUNT;
    const npsl_params &p = c->prm;
    ke = s[SC_KE];  // already the sum over all ranks (canonical order, mesh_kernels.cuh)
    out->kinetic = ke;
    out->ions_kinetic = c->mp.n_ions > 0 ? s[SC_KE_IONS] : 0.0;  // INTERFACE::netIonKE, src/interface.cpp:970-972
    if (c->mp.n_ions > 0) {  // ion-ion energies (the mesh-ion ones are already in the per-vertex sums)
        es += s[SC_ES_IONS];
        lj += s[SC_LJ_IONS];
    }
    out->bending = s[SC_BE];
    out->stretching = s[SC_SE];
    out->total_area = s[SC_AREA];
    out->total_volume = s[SC_VOL];
    out->tension = p.sigma_a * s[SC_AREA];
    const double dv = s[SC_VOL] - p.ref_volume;
    out->volume = p.sigma_v * (dv * dv);
    out->lj = lj;
    out->electrostatic = es;
    out->potential = out->bending + out->stretching + out->tension + out->volume + (out->lj + out->electrostatic);
    out->energy = out->kinetic + out->ions_kinetic + out->potential;
    double mk = 0, mpe = 0, ik = 0, ipe = 0;  // src/newenergies.h:28-46, src/thermostat.h:61-70 (kB = 1)
    for (int j = 0; j < p.chain; j++) {
        mk += 0.5 * th.mesh[j].Q * th.mesh[j].xi * th.mesh[j].xi;
        mpe += th.mesh[j].dof * th.mesh[j].T * th.mesh[j].eta;
        ik += 0.5 * th.ions[j].Q * th.ions[j].xi * th.ions[j].xi;
        ipe += th.ions[j].dof * th.ions[j].T * th.ions[j].eta;
    }
    out->mesh_bath_ke = mk; out->mesh_bath_pe = mpe; out->ions_bath_ke = ik; out->ions_bath_pe = ipe;
    if ((r = xchg_status(c))) return r;
    if (!std::isfinite(out->energy)) return fail(c, NPSL_ERR_NUMERIC, "non-finite energy");
    return NPSL_OK;
}

int npsl_force_calculation(npsl_ctx *c, const double *pos, double *force) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    int r;
    if ((r = upload_records(c, c->xyzq, pos, true, 0))) return r;
    int n = issue(c, false, false, false, false);
    if (n < 0) return n;
    c->launches += n;
    c->forces_valid = true;
    if ((r = download_records(c, c->force, force, true, 1))) return r;
    return NPSL_OK;
}

int npsl_md_run(npsl_ctx *c, const double *pos, const double *vel, const double *q, int n_steps,
                npsl_energies *energies_out, double *pos_out, double *vel_out) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    int r;
    if ((r = npsl_upload_state(c, pos, vel, q))) return r;
    if ((r = npsl_force_init(c))) return r;
    if ((r = npsl_step(c, n_steps))) return r;
    if (energies_out && (r = npsl_energy(c, energies_out))) return r;
    if ((r = npsl_download_state(c, pos_out, vel_out, nullptr))) return r;
    return NPSL_OK;
}

int npsl_md_step_host(npsl_ctx *c, double *pos, double *vel, double *force, int n_steps, double *kinetic_out) {
    if (!c || !pos || !vel || !force) return fail(c, NPSL_ERR_ARG, "NULL argument");
    if (n_steps < 0) return fail(c, NPSL_ERR_ARG, "n_steps < 0");
    int r = ensure_graph(c);
    if (r) return r;
    // host AoS -> pinned staging -> device records: three copies in flight, then the steps, then back
    const size_t bytes = sizeof(double) * 3 * c->nv;
    const int nb = (c->nv + 255) / 256;
    double *in[3] = {pos, vel, force};
    double4 *dev[3] = {c->xyzq, c->vel, c->force};
    CU(cudaStreamSynchronize(c->s_main));
    bool direct[3];
    for (int k = 0; k < 3; k++) {
        direct[k] = user_pinned(c, in[k], bytes);
        const double *src = in[k];
        if (!direct[k]) {
            memcpy(c->h_pin[k], in[k], bytes);
            src = c->h_pin[k];
        }
        CU(cudaMemcpyAsync(c->d_pack[k], src, bytes, cudaMemcpyHostToDevice, c->s_main));
        k_scatter3<<<nb, 256, 0, c->s_main>>>(dev[k], c->d_pack[k], c->nv, k == 0 ? 1 : 0);
    }
    c->launches += 3;
    if ((r = rows_barrier(c))) return r;
    c->forces_valid = true;  // the caller's forvec is the force at the caller's posvec (src/newmd.cpp:114-117)
    c->local_valid = false;
    for (int k = 0; k < n_steps; k++) CU(cudaGraphLaunch(c->step_graph, c->s_main));
    c->launches += (int64_t) c->launches_per_step * n_steps;
    for (int k = 0; k < 3; k++) {
        if (k > 0) {
            r = gather_records(c, dev[k]);
            if (r) return r;
        }
        k_gather3<<<nb, 256, 0, c->s_main>>>(dev[k], c->d_pack[k], c->nv);
        CU(cudaMemcpyAsync(direct[k] ? in[k] : c->h_pin[k], c->d_pack[k], bytes, cudaMemcpyDeviceToHost, c->s_main));
    }
    c->launches += 3;
    if ((r = rows_barrier(c))) return r;
    double ke = 0;
    CU(cudaMemcpyAsync(&ke, c->scal + SC_KE, sizeof(double), cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    for (int k = 0; k < 3; k++)
        if (!direct[k]) memcpy(in[k], c->h_pin[k], bytes);
    if (kinetic_out) *kinetic_out = ke;
    CU(cudaGetLastError());
    return NPSL_OK;
}

int npsl_host_state(npsl_ctx *c, double **pos, double **vel, double **force, double **kinetic) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    int r = ensure_packed_buffers(c);
    if (r) return r;
    const size_t n3 = (size_t) 3 * c->nv;
    if (pos) *pos = c->h_state;
    if (vel) *vel = c->h_state + n3;
    if (force) *force = c->h_state + 2 * n3;
    if (kinetic) *kinetic = c->h_state + 3 * n3;
    return NPSL_OK;
}

int npsl_md_step_packed(npsl_ctx *c, int n_steps) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    if (n_steps < 1) return fail(c, NPSL_ERR_ARG, "n_steps < 1");
    int r = ensure_packed_buffers(c);
    if (r) return r;
    const size_t n3 = (size_t) 3 * c->nv;
    c->forces_valid = true;  // the caller's forvec is the force at the caller's posvec (src/newmd.cpp:114-117)
    c->local_valid = false;
    if (n_steps == 1 && packed_fused_ok(c)) {
        if ((r = ensure_packed_graph(c))) return r;
        CU(cudaGraphLaunch(c->packed_graph, c->s_main));
        c->launches += c->packed_launches;
        CU(cudaStreamSynchronize(c->s_main));
        CU(cudaGetLastError());
        return NPSL_OK;
    }
    // general form: one copy in, one unpacking kernel, the resident steps, one packing kernel, one copy out
    if ((r = ensure_graph(c))) return r;
    const int nb = (c->nv + 255) / 256;
    CU(cudaMemcpyAsync(c->d_state_in, c->h_state, sizeof(double) * 3 * n3, cudaMemcpyHostToDevice, c->s_main));
    k_unpack_state<<<nb, 256, 0, c->s_main>>>(c->xyzq, c->vel, c->force, c->d_state_in, c->nv);
    if ((r = rows_barrier(c))) return r;
    for (int k = 0; k < n_steps; k++) CU(cudaGraphLaunch(c->step_graph, c->s_main));
    if ((r = gather_records(c, c->vel))) return r;
    if ((r = gather_records(c, c->force))) return r;
    k_pack_state<<<nb, 256, 0, c->s_main>>>(c->xyzq, c->vel, c->force, c->scal, c->d_state_out, c->nv);
    if ((r = rows_barrier(c))) return r;
    CU(cudaMemcpyAsync(c->h_state, c->d_state_out, sizeof(double) * (3 * n3 + 1), cudaMemcpyDeviceToHost, c->s_main));
    c->launches += 2 + (int64_t) c->launches_per_step * n_steps;
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    return NPSL_OK;
}

int npsl_sync(npsl_ctx *c) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    int r = xchg_status(c);
    if (r) return r;
    return collect_timing(c);
}

// ---- explicit counterions ---------------------------------------------------------------

int npsl_set_ions(npsl_ctx *c, int n_ions, const double *pos, const double *vel, const double *q, double diameter,
                  double lj_length_mesh_ions) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    if (n_ions < 0 || (n_ions > 0 && (!pos || !q))) return fail(c, NPSL_ERR_ARG, "bad ion arguments");
    if (n_ions > 0 && !(diameter > 0 && lj_length_mesh_ions > 0)) return fail(c, NPSL_ERR_ARG, "ion LJ lengths must be positive");
    CU(cudaStreamSynchronize(c->s_main));
    void *old[] = {c->ion_xyzq, c->ion_vel, c->ion_force, c->ion_mesh, c->ion_e, c->ion_mesh_lj, c->ion_part};
    for (void *p : old)
        if (p) cudaFree(p);
    c->ion_xyzq = c->ion_vel = c->ion_force = c->ion_mesh = nullptr; c->ion_e = nullptr; c->ion_mesh_lj = c->ion_part = nullptr;
    c->mp.n_ions = n_ions;
    drop_graph(c);
    c->forces_valid = c->local_valid = false;
    double zeros[5] = {0, 1.0, 0.5 * c->prm.dt, 0, 0};  // KE, expfac, kick, ES, LJ of the ions until the first force pass
    CU(cudaMemcpy(c->scal + SC_KE_IONS, zeros, sizeof zeros, cudaMemcpyHostToDevice));
    if (n_ions == 0) return NPSL_OK;
    CU(dalloc(&c->ion_xyzq, n_ions)); CU(dalloc(&c->ion_vel, n_ions)); CU(dalloc(&c->ion_force, n_ions));
    const size_t ion_tiles = (n_ions + 127) / 128;
    CU(dalloc(&c->ion_e, n_ions)); CU(dalloc(&c->ion_mesh, ion_tiles * c->nv)); CU(dalloc(&c->ion_mesh_lj, ion_tiles * c->nv));
    CU(dalloc(&c->ion_part, (size_t) ((c->nv + 127) / 128) * 3 * n_ions));
    std::vector<double4> rec(n_ions), vr(n_ions);
    for (int i = 0; i < n_ions; i++) {
        rec[i] = make_double4(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], q[i]);
        vr[i] = vel ? make_double4(vel[3 * i], vel[3 * i + 1], vel[3 * i + 2], 0.0) : make_double4(0, 0, 0, 0);
    }
    CU(cudaMemcpy(c->ion_xyzq, rec.data(), sizeof(double4) * n_ions, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->ion_vel, vr.data(), sizeof(double4) * n_ions, cudaMemcpyHostToDevice));
    IonParams &ip = c->ip;
    ip.n_ions = n_ions; ip.nv = c->nv;
    ip.coef = c->prm.scalefactor / c->prm.em;
    ip.inv_lambda = 1.0 / c->prm.inv_kappa_out;
    ip.elj = c->prm.elj;
    ip.d2_mi = lj_length_mesh_ions * lj_length_mesh_ions; ip.cut2_mi = 1.25992104989487316476 * ip.d2_mi;
    ip.d2_ii = diameter * diameter; ip.cut2_ii = 1.25992104989487316476 * ip.d2_ii;
    ip.dt = c->prm.dt;
    ip.c2s = c->pp.c2s;
    // the reverse half-chain of the first step needs the ions' kinetic energy: same preparation as after npsl_set_thermostat
    k_ion_finish<false><<<1, 256, 0, c->s_main>>>(c->ion_vel, c->ion_force, c->ion_e, c->scal, c->ip);
    k_thermo_prepare<<<1, 32, 0, c->s_main>>>(c->th_cur, c->th_mid, c->scal, c->mp);
    c->launches += 2;
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    return NPSL_OK;
}

int npsl_download_ions(npsl_ctx *c, double *pos, double *vel, double *force) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    const int n = c->mp.n_ions;
    if (n == 0) return NPSL_OK;
    std::vector<double4> tmp(n);
    double *outs[3] = {pos, vel, force};
    double4 *src[3] = {c->ion_xyzq, c->ion_vel, c->ion_force};
    for (int k = 0; k < 3; k++) {
        if (!outs[k]) continue;
        CU(cudaMemcpyAsync(tmp.data(), src[k], sizeof(double4) * n, cudaMemcpyDeviceToHost, c->s_main));
        CU(cudaStreamSynchronize(c->s_main));
        for (int i = 0; i < n; i++) { outs[k][3 * i] = tmp[i].x; outs[k][3 * i + 1] = tmp[i].y; outs[k][3 * i + 2] = tmp[i].z; }
    }
    return NPSL_OK;
}

// ---- data feeds of the writers -------------------------------------------------------

int npsl_vertex_geometry(npsl_ctx *c, double *area, double *normal) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    const VertexTables vt = {c->degree, c->ring, c->ring_l0, c->bslot_ref};
    k_vertex_geometry<<<(c->nv + kMeshThreads - 1) / kMeshThreads, kMeshThreads, 0, c->s_main>>>(c->xyzq, vt, c->vertex_geo, c->mp);
    c->launches++;
    std::vector<double4> tmp(c->nv);
    CU(cudaMemcpyAsync(tmp.data(), c->vertex_geo, sizeof(double4) * c->nv, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    for (int i = 0; i < c->nv; i++) {
        if (area) area[i] = tmp[i].w;
        if (normal) { normal[3 * i] = tmp[i].x; normal[3 * i + 1] = tmp[i].y; normal[3 * i + 2] = tmp[i].z; }
    }
    return NPSL_OK;
}

int npsl_local_energies(npsl_ctx *c, double *electrostatic, double *elastic, double *bending, double *stretching) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    if (!c->forces_valid) return fail(c, NPSL_ERR_STATE, "npsl_local_energies before npsl_force_init");
    if (!c->local_valid) {  // an energy-enabled, storing force pass at the current positions
        int r = force_pass(c);
        if (r) return r;
    }
    if (c->world > 1 && !c->replicated) {  // row-sharded: the per-vertex energies of the other ranks' rows
        NC(g_nccl.AllGather(c->vertex_e + (size_t) c->rank * c->range, c->vertex_e, (size_t) c->range * 2, ncclFloat64_, c->comm,
                            c->s_main));
    }
    k_local_faces<<<c->n_face_blocks, kMeshThreads, 0, c->s_main>>>(c->face_v, c->face_e, c->vertex_e, c->edge_E, c->face_local,
                                                                    c->mp);
    c->launches++;
    std::vector<double4> tmp(c->nf);
    CU(cudaMemcpyAsync(tmp.data(), c->face_local, sizeof(double4) * c->nf, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    CU(cudaGetLastError());
    for (int f = 0; f < c->nf; f++) {
        if (electrostatic) electrostatic[f] = tmp[f].x;
        if (elastic) elastic[f] = tmp[f].y;
        if (bending) bending[f] = tmp[f].z;
        if (stretching) stretching[f] = tmp[f].w;
    }
    return NPSL_OK;
}

// ---- introspection -----------------------------------------------------------------------------------

int64_t npsl_launch_count(const npsl_ctx *c) { return c ? c->launches : 0; }

int npsl_pair_timing(npsl_ctx *c, int enable) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    int r = collect_timing(c);
    if (r) return r;
    c->timing = enable != 0;
    if (c->timing && c->t_ev.empty()) {
        c->t_ev.resize(512);
        for (auto &e : c->t_ev) CU(cudaEventCreate(&e));
    }
    c->t_sum_ms = 0;
    c->t_count = 0;
    return NPSL_OK;
}

int npsl_pair_timing_read(npsl_ctx *c, double *mean_ms, int64_t *launches) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    int r = collect_timing(c);
    if (r) return r;
    if (mean_ms) *mean_ms = c->t_count ? c->t_sum_ms / (double) c->t_count : 0.0;
    if (launches) *launches = c->t_count;
    return NPSL_OK;
}

// Shared by npsl_step_timing and npsl_step_timing_list: lead_in untimed steps, then n_steps timed ones, all submitted
// to the stream in one go. Buffers and events are created before anything is enqueued, so that no allocation of this
// rank can land inside a peer's bracket (a rank's step waits for every peer's exchange signal).
static int step_timing_impl(npsl_ctx *c, int n_steps, int flush_l2, int lead_in, double *per_step_ms, double *total_ms) {
    if (n_steps < 1 || lead_in < 0) return fail(c, NPSL_ERR_ARG, "n_steps < 1 or lead_in < 0");
    if (!c->forces_valid) return fail(c, NPSL_ERR_STATE, "npsl_step_timing before npsl_force_init");
    c->local_valid = false;
    int r = ensure_graph(c);
    if (r) return r;
    const size_t need = flush_l2 ? (size_t) 2 * n_steps : 2;
    while (c->s_ev.size() < need) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        c->s_ev.push_back(e);
    }
    if (flush_l2 && !c->flush_buf) {
        c->flush_bytes = (size_t) 256 << 20;  // 2x the 126 MB L2
        CU(cudaMalloc(&c->flush_buf, c->flush_bytes));
        CU(cudaMemset(c->flush_buf, 0, c->flush_bytes));  // touched once here, never inside a timed call
        CU(cudaDeviceSynchronize());
    }
    for (int k = 0; k < lead_in; k++) {
        if (flush_l2) CU(cudaMemsetAsync(c->flush_buf, k & 0xff, c->flush_bytes, c->s_main));
        CU(cudaGraphLaunch(c->step_graph, c->s_main));
    }
    double sum = 0;
    if (!flush_l2) {
        CU(cudaEventRecord(c->s_ev[0], c->s_main));
        for (int k = 0; k < n_steps; k++) CU(cudaGraphLaunch(c->step_graph, c->s_main));
        CU(cudaEventRecord(c->s_ev[1], c->s_main));
        CU(cudaEventSynchronize(c->s_ev[1]));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, c->s_ev[0], c->s_ev[1]));
        sum = ms;
        if (per_step_ms)
            for (int k = 0; k < n_steps; k++) per_step_ms[k] = ms / n_steps;
    } else {
        // per-step brackets; the flush sits between them
        for (int k = 0; k < n_steps; k++) {
            CU(cudaMemsetAsync(c->flush_buf, k & 0xff, c->flush_bytes, c->s_main));
            CU(cudaEventRecord(c->s_ev[2 * k], c->s_main));
            CU(cudaGraphLaunch(c->step_graph, c->s_main));
            CU(cudaEventRecord(c->s_ev[2 * k + 1], c->s_main));
        }
        CU(cudaStreamSynchronize(c->s_main));
        for (int k = 0; k < n_steps; k++) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, c->s_ev[2 * k], c->s_ev[2 * k + 1]));
            sum += ms;
            if (per_step_ms) per_step_ms[k] = ms;
        }
    }
    c->launches += (int64_t) c->launches_per_step * (n_steps + lead_in);
    CU(cudaGetLastError());
    if (total_ms) *total_ms = sum;
    return NPSL_OK;
}

int npsl_step_timing(npsl_ctx *c, int n_steps, int flush_l2, double *total_ms) {
    if (!c || !total_ms) return fail(c, NPSL_ERR_ARG, "NULL argument");
    return step_timing_impl(c, n_steps, flush_l2, 0, nullptr, total_ms);
}

int npsl_step_timing_list(npsl_ctx *c, int n_steps, int flush_l2, int lead_in, double *per_step_ms) {
    if (!c || !per_step_ms) return fail(c, NPSL_ERR_ARG, "NULL argument");
    return step_timing_impl(c, n_steps, flush_l2, lead_in, per_step_ms, nullptr);
}

int npsl_step_profile(npsl_ctx *c, int n_steps, char *out, int out_len) {
    if (!c || !out || out_len < 64) return fail(c, NPSL_ERR_ARG, "bad argument");
    if (!c->forces_valid) return fail(c, NPSL_ERR_STATE, "npsl_step_profile before npsl_force_init");
    c->local_valid = false;
    std::vector<double> sum;
    std::vector<const char *> names;
    for (int k = 0; k < n_steps; k++) {
        c->prof = true;
        c->prof_used = 0;
        int n = issue(c, true, false, false, false);
        c->prof = false;
        if (n < 0) return n;
        c->launches += n;
        CU(cudaStreamSynchronize(c->s_main));
        if (sum.empty()) { sum.assign(c->prof_used, 0.0); names.assign(c->prof_names.begin(), c->prof_names.begin() + c->prof_used); }
        for (size_t j = 1; j < c->prof_used && j < sum.size(); j++) {
            float ms = 0;
            CU(cudaEventElapsedTime(&ms, c->prof_ev[j - 1], c->prof_ev[j]));
            sum[j] += ms;
        }
    }
    std::string txt;
    char line[160];
    double tot = 0;
    for (size_t j = 1; j < sum.size(); j++) {
        snprintf(line, sizeof line, "%-34s %9.2f us\n", names[j], 1e3 * sum[j] / n_steps);
        txt += line;
        tot += sum[j];
    }
    snprintf(line, sizeof line, "%-34s %9.2f us (direct launches, not the graph)\n", "step", 1e3 * tot / n_steps);
    txt += line;
    snprintf(out, out_len, "%s", txt.c_str());
    return NPSL_OK;
}

int npsl_set_pair_plan(npsl_ctx *c, int rows_per_thread, int col_splits) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    const bool sym_only = rows_per_thread >= 5 && rows_per_thread <= 8;  // shapes of the symmetric kernel only
    if (!(rows_per_thread == 0 || rows_per_thread == 1 || rows_per_thread == 2 || rows_per_thread == 4 || (sym_only && c->use_sym)) ||
        col_splits < 0)
        return fail(c, NPSL_ERR_ARG, "rows_per_thread must be 0,1,2,4 (6,8: symmetric kernel) and col_splits >= 0");
    CU(cudaStreamSynchronize(c->s_main));
    c->plan_rows = rows_per_thread;
    c->plan_splits = col_splits;
    if (c->use_sym) {  // measurement knob for the symmetric kernel: (rows per thread, threads per CTA [+1: tighter register cap])
        if (rows_per_thread) c->sym_rows = rows_per_thread;
        c->sym_shape_set = true;
        if (col_splits >= 32) { c->sym_threads = col_splits & ~1; c->sym_minb = (col_splits & 1) ? 8 : 0; }
        c->plan_rows = c->plan_splits = 0;
        plan_pair(c, c->n_sm);
        int r = plan_sym(c);
        if (r) return r;
    }
    plan_pair(c, c->n_sm);
    if (c->pair_partial) cudaFree(c->pair_partial);
    c->pair_partial = nullptr;
    CU(dalloc(&c->pair_partial, (size_t) c->pair_splits * c->pair_row_stride));
    apply_pair_plan(c);
    drop_graph(c);
    c->forces_valid = false;
    return NPSL_OK;
}

int npsl_set_volume_constraint(npsl_ctx *c, int on) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    CU(cudaStreamSynchronize(c->s_main));
    c->mp.constraint = on ? 1 : 0;
    drop_graph(c);
    c->forces_valid = false;  // neg_GradVi has to be (re)computed by a force evaluation
    return NPSL_OK;
}

int npsl_constraint_init(npsl_ctx *c) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    if (!c->mp.constraint) return fail(c, NPSL_ERR_STATE, "npsl_constraint_init without npsl_set_volume_constraint");
    if (!c->forces_valid) return fail(c, NPSL_ERR_STATE, "npsl_constraint_init before npsl_force_init");
    int r;
    if ((r = gather_records(c, c->vel))) return r;  // rows plan, see issue(); neg_GradVi was gathered by the force pass
    k_constraint_sums<0><<<c->n_face_blocks, kMeshThreads, 0, c->s_main>>>(c->xyzq, c->vel, c->ngv, c->face_v, c->part_con,
                                                                          c->tickets + 2, c->scal, c->mp);
    k_shake_apply<<<(c->nv + kMeshThreads - 1) / kMeshThreads, kMeshThreads, 0, c->s_main>>>(c->xyzq, c->vel, c->ngv, c->scal,
                                                                                           c->mp);
    k_constraint_sums<1><<<c->n_face_blocks, kMeshThreads, 0, c->s_main>>>(c->xyzq, c->vel, c->ngv, c->face_v, c->part_con,
                                                                          c->tickets + 2, c->scal, c->mp);
    k_rattle_apply<true><<<c->n_vertex_blocks, kMeshThreads, 0, c->s_main>>>(c->vel, c->ngv, c->part_vert, c->tickets + 1,
                                                                             c->scal, c->ke_warp, c->n_ke_warps, c->th_mid,
                                                                             c->th_cur, c->d_peer_ke, c->xg, c->mp);
    c->launches += 4;
    if (c->world > 1 && !c->replicated) {  // rows plan: the kinetic energy is finished after the exchange of the partials
        const size_t per_rank = (size_t) c->range / 32;
        if (!c->p2p)
            NC(g_nccl.AllGather(c->ke_warp + (size_t) c->rank * per_rank, c->ke_warp, per_rank, ncclFloat64_, c->comm, c->s_main));
        k_ke_post<<<1, kMeshThreads, 0, c->s_main>>>(c->ke_warp, c->n_ke_warps, c->scal, c->th_mid, c->th_cur, 0, c->xg, c->mp);
        c->launches++;
    }
    CU(cudaGetLastError());
    return NPSL_OK;
}

int npsl_constraint_multipliers(npsl_ctx *c, double *shake_lambda, double *rattle_mu) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    double v[2];
    CU(cudaMemcpyAsync(v, c->scal + SC_LAMBDA, sizeof v, cudaMemcpyDeviceToHost, c->s_main));
    CU(cudaStreamSynchronize(c->s_main));
    if (shake_lambda) *shake_lambda = v[0];
    if (rattle_mu) *rattle_mu = v[1];
    return NPSL_OK;
}

int npsl_set_pair_mode(npsl_ctx *c, int mode) {
    if (!c) return fail(c, NPSL_ERR_ARG, "ctx is NULL");
    if (mode < 0 || mode > 2) return fail(c, NPSL_ERR_ARG, "mode must be 0 (automatic), 1 (ordered) or 2 (symmetric)");
    CU(cudaStreamSynchronize(c->s_main));
    c->pair_mode = mode;
    int r = plan_sym(c);
    if (r) return r;
    if (mode == 2 && !c->use_sym) return fail(c, NPSL_ERR_ARG, "symmetric pair kernel needs a rank count that divides 8");
    if (c->world > 1 && (r = apply_shard_plan(c))) return r;
    drop_graph(c);
    c->forces_valid = false;
    return NPSL_OK;
}

int npsl_pair_mode(const npsl_ctx *c) { return c ? (c->use_sym ? 2 : 1) : NPSL_ERR_ARG; }

int npsl_exchange_mode(const npsl_ctx *c) { return c ? (c->world == 1 ? 0 : (c->p2p ? 2 : 1)) : NPSL_ERR_ARG; }

int npsl_plan_pair_units(int n_vertices, int rank, int world, int32_t *units, int max_units, int32_t *geometry) {
    npsl_ctx *c = nullptr;
    if (n_vertices < 1 || world < 1 || rank < 0 || rank >= world || kSymV % world != 0)
        return fail(c, NPSL_ERR_ARG, "bad arguments (the rank count must divide 8)");
    SymParams sp;
    memset(&sp, 0, sizeof sp);
    sym_geometry(n_vertices, 6 * 64, sp);  // the default shape of k_pair_sym
    std::vector<SymUnit> mine;
    const int v_count = kSymV / world;
    sym_units_of(sp, rank * v_count, v_count, mine);
    if (geometry) {
        geometry[0] = sp.rb; geometry[1] = kSymCT; geometry[2] = sp.n_chunks;
        geometry[3] = sp.n_coarse; geometry[4] = sp.chunk_tiles; geometry[5] = sp.fine_tiles;
    }
    for (int k = 0; k < (int) mine.size() && k < max_units; k++) {
        const int t0 = std::max(sym_chunk_tile0(sp, mine[k].cc) * kSymCT, mine[k].I * sp.rb);
        const int t1 = std::min(sym_chunk_tile0(sp, mine[k].cc + 1) * kSymCT, n_vertices);
        units[4 * k] = mine[k].I; units[4 * k + 1] = mine[k].cc; units[4 * k + 2] = t0; units[4 * k + 3] = t1;
    }
    return (int) mine.size();
}

int npsl_shard_plan(const npsl_ctx *c) { return c ? (c->world == 1 ? 0 : (c->replicated ? 2 : 1)) : NPSL_ERR_ARG; }

int npsl_owned_rows(const npsl_ctx *c, int32_t *lo, int32_t *hi) {
    if (!c) return NPSL_ERR_ARG;
    if (lo) *lo = c->row_lo;
    if (hi) *hi = c->row_hi;
    return NPSL_OK;
}

int npsl_pair_geometry(const npsl_ctx *c, int32_t *rows_per_thread, int32_t *row_tiles, int32_t *col_splits,
                       int32_t *cols_per_cta) {
    if (!c) return NPSL_ERR_ARG;
    if (rows_per_thread) *rows_per_thread = c->pair_rows_per_thread;
    if (row_tiles) *row_tiles = c->pair_row_tiles;
    if (col_splits) *col_splits = c->pair_splits;
    if (cols_per_cta) *cols_per_cta = c->pair_cols_per_cta;
    return NPSL_OK;
}

int npsl_pair_poly_eval(double inv_kappa_out, const double *s, double *table_value, double *exact_value, int n) {
    npsl_ctx *c = nullptr;
    if (!(inv_kappa_out > 0) || !s || n < 0) return fail(c, NPSL_ERR_ARG, "bad argument");
    std::vector<double> tab;
    build_pair_table(inv_kappa_out, tab);
    for (int i = 0; i < n; i++) {
        if (table_value) table_value[i] = eval_pair_table(tab, s[i]);
        if (exact_value) exact_value[i] = (double) radial_factor((long double) s[i], (long double) inv_kappa_out);
    }
    return NPSL_OK;
}

// The default radial-factor evaluation of k_pair_sym (SYM_FAST, pair_sym_kernel.cuh), operation for operation on the host:
// s = r^2 in units of lambda^2. The device's rsqrt seed comes from MUFU.RSQ64H (not available here): it is modelled as
// 1/sqrt(s) cut to seed_bits mantissa bits, so the test can bound the scheme's error for any seed at least that good.
int npsl_pair_fast_eval(const double *s, double *value, double *exact_value, int n, int seed_bits) {
    npsl_ctx *c = nullptr;
    if (!s || n < 0 || seed_bits < 8 || seed_bits > 52) return fail(c, NPSL_ERR_ARG, "bad argument");
    static const double tab[kExpTabFast] = {NPSL_EXP2FAST_TABLE};
    const double MAGIC = 6755399441055744.0, c2p = 1.4426950408889634074 * (double) kExpTabFast;
    for (int i = 0; i < n; i++) {
        const double a = s[i];
        if (value) {
            double y0 = (double) (1.0L / sqrtl((long double) a));
            uint64_t yb;
            memcpy(&yb, &y0, 8);
            yb &= ~((uint64_t(1) << (52 - seed_bits)) - 1);
            memcpy(&y0, &yb, 8);
            const double b = a * y0;
            const double e = fma(-b, y0, 1.0);
            const double y = fma(y0 * e, 0.5, y0);
            const double rn = -(a * y);
            const double tt = fma(rn, c2p, MAGIC);
            const double f = fma(rn, c2p, -(tt - MAGIC));
            uint64_t tb;
            memcpy(&tb, &tt, 8);
            const int32_t nn = (int32_t) (uint32_t) tb;  // low word: round(-r log2(e) 2^TB)
            const int j = nn & (kExpTabFast - 1);
            uint64_t Tb;
            memcpy(&Tb, &tab[j], 8);
            Tb -= (uint64_t) j << (32 + 20 - NPSL_EXP2FAST_TB);            // the staged table (plan_sym)
            const uint32_t hi = (uint32_t) (Tb >> 32) + (uint32_t) (nn * (1 << (20 - NPSL_EXP2FAST_TB)));  // one multiply-add
            Tb = ((uint64_t) hi << 32) | (Tb & 0xffffffffu);
            double T;
            memcpy(&T, &Tb, 8);
            const double g = fma(f, NPSL_EXP2FAST_G2, NPSL_EXP2FAST_G1);
            const double ex = fma(T, f * g, T);
            const double y2 = y * y;
            value[i] = ex * fma(y2, y, y2);
        }
        if (exact_value) {
            const long double r = sqrtl((long double) a);
            exact_value[i] = (double) (expl(-r) * (1.0L / (r * r * r) + 1.0L / (r * r)));
        }
    }
    return NPSL_OK;
}

int npsl_pair_kernel_info(const npsl_ctx *c, int32_t *rows_per_thread, int32_t *threads, double *fp64_instr_per_pair) {
    if (!c) return NPSL_ERR_ARG;
    if (!c->use_sym) {  // ordered kernel: SASS count of k_pair's inner loop, per ordered pair (pair_kernel.cuh)
        if (rows_per_thread) *rows_per_thread = c->pair_rows_per_thread;
        if (threads) *threads = kPairThreads;
        if (fp64_instr_per_pair) *fp64_instr_per_pair = 27.0;
        return NPSL_OK;
    }
    if (rows_per_thread) *rows_per_thread = c->sym_rows;
    if (threads) *threads = c->sym_threads;
    // per UNORDERED pair, from the SASS of the inner loops (tools/sass_loops.py): 3 sub + 3 r^2 + the radial factor
    // (17 exact / 14 reduced accuracy / 6 table) + 2 charge products + 6 accumulations (the column accumulators travel
    // through registers, no separate adds)
    const double radial = c->sym_math == SYM_EXACT ? 17.0 : (c->sym_math == SYM_FAST ? 14.0 : 6.0);
    if (fp64_instr_per_pair) *fp64_instr_per_pair = 14.0 + radial;
    return NPSL_OK;
}

int npsl_pair_math(const npsl_ctx *c) {
    if (!c) return NPSL_ERR_ARG;
    return c->use_sym ? c->sym_math : SYM_EXACT;
}

int npsl_fp64_peak(double *dfma_per_s, double *sm_clock_mhz) {
    npsl_ctx *c = nullptr;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    double *out = nullptr;
    CU(cudaMalloc((void **) &out, sizeof(double)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    k_dfma_peak<<<blocks, threads>>>(out, 1 << 10, 1.0000001, 0.9999999);  // warm-up
    double best = 0;
    for (int rep = 0; rep < 3; rep++) {
        CU(cudaEventRecord(e0));
        k_dfma_peak<<<blocks, threads>>>(out, iters, 1.0000001, 0.9999999);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double) blocks * threads * 16.0 * iters / (ms * 1e-3);
        best = std::max(best, rate);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    CU(cudaGetLastError());
    if (dfma_per_s) *dfma_per_s = best;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    if (sm_clock_mhz) *sm_clock_mhz = khz / 1000.0;
    return NPSL_OK;
}

}  // extern "C"
