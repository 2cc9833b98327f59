// Internal declarations of libvfvmb200.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/vfvm_b200.h"
#include "peer.cuh"

#define VFVM_MAX_SPECIES 10
#define VFVM_MAX_PARAMS 160
#define VFVM_MAX_BC 32
#define VFVM_MAX_BREGIONS 16
#define VFVM_MAX_CREGIONS 32

// ------------------------------------------------------------------------------------------------ device buffers
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int64_t* tally = nullptr;  // handle-wide byte counter
    void alloc(size_t count) {
        if (count == n && p) return;
        release();
        n = count;
        if (count) {
            cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
            if (e != cudaSuccess) {
                p = nullptr;
                n = 0;
                throw std::string("cudaMalloc of ") + std::to_string(count * sizeof(T)) + " bytes failed: " + cudaGetErrorString(e);
            }
            if (tally) *tally += (int64_t)(count * sizeof(T));
        }
    }
    void release() {
        if (p) {
            cudaFree(p);
            if (tally) *tally -= (int64_t)(n * sizeof(T));
        }
        p = nullptr;
        n = 0;
    }
    void upload(const T* h, size_t count, cudaStream_t s) {
        alloc(count);
        if (count) cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void download(T* h, cudaStream_t s) const {
        if (n) cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
    }
    std::vector<T> to_host(cudaStream_t s) const {
        std::vector<T> v(n);
        download(v.data(), s);
        return v;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

// ------------------------------------------------------------------------------------------------ physics block (kernel argument)
struct PhysSlotDev {
    int id;
    int np;
    int off;  // offset into params[]
};
struct PhysicsDev {
    PhysSlotDev slot[VFVM_NUM_SLOTS];
    double params[VFVM_MAX_PARAMS];
    int nbc;
    vfvm_bc_entry bc[VFVM_MAX_BC];
    int has_legacy_bc;
    int nbregions;
    double bfactors[VFVM_MAX_SPECIES * VFVM_MAX_BREGIONS];  // n x nbregions, column-major
    double bvalues[VFVM_MAX_SPECIES * VFVM_MAX_BREGIONS];
    const double* nodal_source;  // n x N or null
};

// species-coupling masks: bit (i*n + j) set <=> d f_i / d u_j is not identically zero
struct Masks {
    uint64_t flux[2];      // we support n <= 10 -> 100 bits
    uint64_t reaction[2];
    uint64_t storage[2];
    uint64_t boundary[2];  // union over all boundary contributions
    uint64_t edgereaction[2];
};
static inline bool mask_get(const uint64_t* m, int bit) { return (m[bit >> 6] >> (bit & 63)) & 1ull; }
static inline void mask_set(uint64_t* m, int bit) { m[bit >> 6] |= (1ull << (bit & 63)); }

// halo description of a coarser AMG level: same neighbour ranks as level 0 (handle), shorter lists
struct LevelHalo {
    int64_t Nown = 0, nhalo = 0;
    std::vector<int64_t> send_ptr, recv_ptr;  // per neighbour slot, nn+1 entries
    DevBuf<int32_t> send_idx;                 // owned nodes of this level to send, grouped by neighbour
    DevBuf<double> send_buf;                  // NCCL transport: packed values
};

// chunking of the pipelined host-vector assembly (vfvm_eval_res_jac with VFVM_HOST)
struct PipePlan {
    bool valid = false;
    int K = 0;
    std::vector<int> slice_begin;   // K+1: chunk c = slices [slice_begin[c], slice_begin[c+1])
    std::vector<int> piece_hi;      // K: last piece of U the (owned) columns of chunk c reach
    std::vector<int> needs_halo;    // K: chunk c has halo columns (several ranks): it also waits for the halo piece
    std::vector<int64_t> bn_begin;  // K+1: boundary nodes of chunk c
};

// ------------------------------------------------------------------------------------------------ handle
struct vfvm_handle {
    int device = 0;
    cudaStream_t stream = nullptr, stream_in = nullptr, stream_out = nullptr;  // main stream; copy streams of the pipelined host path
    std::vector<cudaEvent_t> pipe_ev;
    PipePlan pipe;
    int grid_pct = 100;  // share of the persistent grid the row kernels may use (pipelined host path: leaves HBM slots to the copy engines)
    std::vector<int32_t> bn_node_host;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr, ev4 = nullptr;
    std::string err;
    int64_t bytes = 0;
    int64_t launches = 0;
    double times[VFVM_NUM_TIMES] = {0};

    // grid
    int dim = 0, coordsys = 0;
    int64_t N = 0, C = 0, NB = 0, Nown = 0;
    int ncellregions = 0, nbfaceregions = 0;
    DevBuf<double> coord;
    DevBuf<int32_t> cellnodes, cellregions, bfacenodes, bfaceregions;
    bool have_grid = false, have_geometry = false, have_system = false, have_pattern = false;

    // geometry (K1, K2)
    int64_t E = 0;
    DevBuf<int32_t> edgenodes, celledges;
    DevBuf<int64_t> nf_colptr, ef_colptr;
    DevBuf<int32_t> nf_region, ef_region;
    DevBuf<double> nf_fac, ef_fac;
    DevBuf<double> bfacenodefac;
    bool single_region = true;
    int the_region = 1;

    // system
    int n = 0;
    std::vector<uint8_t> region_species;  // n x ncellregions: species enabled per cell region
    std::vector<uint8_t> bregion_species;  // n x nbregions_bs: boundary species (enable_boundary_species!), may be empty
    int nbregions_bs = 0;
    bool masked = false;                   // some species is not enabled in every cell region
    std::vector<int32_t> node_active_host;
    DevBuf<int32_t> node_active;           // masked systems: bit i set <=> species i is defined at the node (node_dof, src/vfvm_system.jl:445-456)
    PhysicsDev phys;
    DevBuf<PhysicsDev> phys_dev;  // device copy, refreshed by vfvm_sync_physics
    bool phys_dirty = true;
    bool asm_pending = false;  // an assembly was enqueued whose status / timings have not been collected yet
    Masks masks;
    DevBuf<double> nodal_source;
    std::vector<double> host_params;

    // pattern (K3): off-diagonal block CSR over owned rows + separate diagonal blocks
    int64_t nnz_off = 0, nnz_sell = 0;  // true / padded number of off-diagonal blocks
    DevBuf<int32_t> rowptr, colidx;  // rowptr: Nown+1 (CSR offsets = row lengths); colidx: nnz_sell in SELL-32 order
    DevBuf<int32_t> sell_ptr;        // nslices+1: first entry of each 32-row slice
    DevBuf<int32_t> nz_edge;         // nnz -> edge id
    DevBuf<double> nzfac;            // single-region fast path: edge factor per nnz
    DevBuf<int32_t> itemptr;         // multi-region: nnz -> [itemptr[k], itemptr[k+1]) into ef_region / ef_fac
    DevBuf<int32_t> tile_row;        // row tiles of the streaming kernels
    int ntiles = 0, tile_nnz = 0;
    int group_R = 16, ngroups = 0, group_maxnnz = 0;  // warp row groups of the assembly / SpMV kernels
    DevBuf<double> src_cache;                         // tabulated source callback, n x N
    DevBuf<double> node_q;                            // node-transformed unknowns q(u) of flux_node_transform fluxes, n x N
    // boundary nodes: CSR node -> (bface, local node) in bface order
    int64_t nbnodes = 0, nbitems = 0;
    DevBuf<int32_t> bn_node, bn_ptr, bn_bface, bn_local;
    // plane tables
    int cF = 0, cD = 0;           // number of stored planes in off-diagonal / diagonal blocks
    int planeF[100], planeD[100]; // plane -> i*n+j
    int idxF[100], idxD[100];     // i*n+j -> plane or -1
    bool seen_transient = false;
    std::vector<uint8_t> bnode_mask_host;  // per boundary node: extra diag bits (for the exported scalar pattern)

    // values
    DevBuf<float> offval32;  // fp32 copy of offval for the SpMVs inside the AMG cycle (amg.cu: Amg::fp32), refreshed by the numeric setup
    DevBuf<double> offval;   // cF planes x nnz_off
    DevBuf<double> diagval;  // cD planes x Nown
    DevBuf<double> vec[4];   // SOLUTION, OLDSOL, RESIDUAL, UPDATE: n x N (N incl. halo)
    DevBuf<int32_t> flags;   // [0]: NaN seen
    int32_t* flags_host = nullptr;  // pinned

    // linear solver
    int krylov = VFVM_KRYLOV_BICGSTAB, precon = VFVM_PRECON_JACOBI, gmres_restart = 30;
    bool precon_valid = false;
    int last_converged = 0;      // outcome of the last vfvm_linsolve (vfvm_linsolve_status)
    double last_rhsnorm = 0.0;
    DevBuf<double> work[12];
    DevBuf<double> pc_diag;  // (block-)Jacobi inverse blocks, n*n planes x Nown
    DevBuf<double> ilu_off, ilu_diag;
    DevBuf<int32_t> ilu_rank, ilu_lrows, ilu_urows;  // elimination order; rows sorted by lower / upper dependency level
    std::vector<int32_t> ilu_lptr, ilu_uptr;         // level boundaries (host)
    bool ilu_struct_valid = false;
    int ilu_order = 0;
    DevBuf<int32_t> upos;  // first off-diagonal entry with col > row
    DevBuf<double> red;    // reduction scratch
    double* red_host = nullptr;  // pinned

    // comm
    void* nccl = nullptr;  // ncclComm_t
    int rank = 0, nranks = 1;
    std::vector<int32_t> nb_ranks;
    std::vector<int64_t> send_ptr, recv_ptr;
    DevBuf<int32_t> send_idx;
    DevBuf<double> send_buf;
    // peer memory (CUDA IPC mailboxes, peer.cuh)
    bool peer_ok = false;
    char* peer_box = nullptr;                 // my mailbox
    std::vector<char*> peer_base;             // every rank's mailbox as mapped here ([rank] = peer_box)
    std::vector<int64_t> peer_recv_off, peer_slot, peer_halo_doubles;  // per neighbour slot: my place in that neighbour's mailbox
    DevBuf<unsigned long long> peer_seq;      // device-resident sequence counters: [0] halo exchanges, [1] reductions
    long long peer_timeout_ns = 30000000000ll;  // bound of a wait on a peer (VFVM_PEER_TIMEOUT_MS)
    DevBuf<unsigned int> peer_counter;
    // all-gather box of the replicated AMG levels (comm.cu: vfvm_gather_box_create): a second IPC allocation, mapped by every rank
    char* gbox = nullptr;
    std::vector<char*> gbox_base;             // every rank's box as mapped here ([rank] = gbox)
    int64_t gbox_cap = 0;                     // doubles per rank segment
    bool gbox_ok = false;
    DevBuf<unsigned long long> gbox_seq;      // device-resident sequence counter of the all-gathers
    DevBuf<unsigned int> gbox_count;
    void* amg = nullptr;  // aggregation AMG hierarchy (amg.cu)
    // one Krylov iteration captured as a CUDA graph (linsolve.cu): valid while its signature (method, buffers, graph_epoch) holds
    void* iter_graph = nullptr;  // cudaGraphExec_t
    uint64_t iter_graph_sig = 0;
    int64_t iter_graph_launches = 0;
    uint64_t graph_epoch = 0;  // bumped by everything that changes what a captured iteration bakes in (solver options, AMG hierarchy / options)
    bool in_capture = false;   // the handle's stream is being captured: preconditioners must enqueue plain kernels
    bool amg_nccl_in_cycle = false;  // the AMG cycle contains a host-enqueued NCCL collective (replicated levels without a gather box): iterations stay eager
};

#define VFVM_TRY(h, ...)                                        \
    try {                                                       \
        __VA_ARGS__                                             \
    } catch (const std::string& e) {                            \
        (h)->err = e;                                           \
        return VFVM_ERR_CUDA;                                   \
    } catch (const std::exception& e) {                         \
        (h)->err = e.what();                                    \
        return VFVM_ERR_CUDA;                                   \
    }

static inline int vfvm_fail(vfvm_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}
static inline void vfvm_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::string(what) + ": " + cudaGetErrorString(e);
}
#define CK(x) vfvm_check((x), #x)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// one matrix in the DBSR / SELL-32 layout + the vectors of one SpMV (kernel argument)
struct SpmvArgs {
    const int32_t* __restrict__ sell_ptr;
    const int32_t* __restrict__ colidx;
    const double* __restrict__ offval;
    const float* __restrict__ offval32;  // optional fp32 copy of the off-diagonal planes (preconditioner-internal products only)
    const double* __restrict__ diagval;
    const double* __restrict__ x;
    double* __restrict__ y;
    const double* __restrict__ w;  // optional: fused dots (y,w) and (y,y)
    double* __restrict__ part;     // 2 x gridDim partial sums
    int64_t nnz_sell, Nown;
    int nslices;
    signed char idxF[100], idxD[100];
};

// implemented across the .cu files
int vfvm_geometry_build(vfvm_handle* h);
int vfvm_pattern_build(vfvm_handle* h);
int vfvm_assemble_impl(vfvm_handle* h, double time, double tstep, double lambda, bool async = false);
int vfvm_assemble_finish(vfvm_handle* h);
bool vfvm_pipeline_applies(vfvm_handle* h);
int vfvm_eval_res_jac_pipelined(vfvm_handle* h, const double* U, const double* UOld, double* F, double time, double tstep, double lambda);
int vfvm_init_dirichlet_impl(vfvm_handle* h, double time, double lambda);
int vfvm_physics_masks(vfvm_handle* h);
void vfvm_spmv_impl(vfvm_handle* h, const double* x, double* y, bool planes32 = false);
void vfvm_sync_physics(vfvm_handle* h);
void vfvm_source_cache(vfvm_handle* h);
void vfvm_zero_inactive(vfvm_handle* h, double* vec);
SpmvArgs vfvm_spmv_args(vfvm_handle* h);
void vfvm_spmv_level(vfvm_handle* h, SpmvArgs a, const double* x, double* y);
void vfvm_blockinv_level(vfvm_handle* h, const SpmvArgs& a, int64_t N, const double* diagval, double* inv);
void vfvm_amg_setup(vfvm_handle* h);
void vfvm_amg_apply(vfvm_handle* h, const double* in, double* out);
void vfvm_amg_free(vfvm_handle* h);
int vfvm_halo_exchange_level(vfvm_handle* h, LevelHalo& c, double* x);
PeerArgs vfvm_peer_args_halo_level(vfvm_handle* h, const LevelHalo& c);
void vfvm_spmv_level_halo(vfvm_handle* h, SpmvArgs a, LevelHalo& lh, double* x, double* y);
int vfvm_halo_exchange_ptr(vfvm_handle* h, double* x);
int vfvm_comm_allreduce_sum(vfvm_handle* h, double* dev, int count);
int vfvm_comm_allreduce_max(vfvm_handle* h, double* dev, int count);
int vfvm_comm_allgather(vfvm_handle* h, const double* send, double* recv, int64_t count);
int vfvm_comm_allgather_bytes(vfvm_handle* h, void* buf, size_t bytes_per_rank);  // in place: rank r's piece at buf + r * bytes_per_rank
void vfvm_comm_group_start();
void vfvm_comm_group_end();
// all-gather of vec = [nranks][cap] doubles in place (rank r owns segment r): one kernel over the gather box when the peers are mapped,
// ncclAllGather otherwise
int vfvm_gather_box_create(vfvm_handle* h, int64_t cap_doubles);
void vfvm_gather_box_free(vfvm_handle* h);
int vfvm_allgather_segments(vfvm_handle* h, double* vec, int64_t cap);
int vfvm_peer_check(vfvm_handle* h);  // after peer kernels: VFVM_ERR_COMM if a wait on a peer timed out
PeerArgs vfvm_peer_args_halo(vfvm_handle* h);    // arguments of a halo exchange (the sequence number lives on the device)
PeerArgs vfvm_peer_args_reduce(vfvm_handle* h);  // arguments of a reduction

