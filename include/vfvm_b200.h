/*
 * vfvm_b200.h -- C ABI of libvfvmb200.so, the B200 (sm_100a) Newton hot path for VoronoiFVM.jl.
 *
 * The reference (VoronoiFVM.jl v3.5.2, pure Julia) has no FFI boundary of its own; the plug points this
 * ABI sits behind are Julia method dispatch on
 *     eval_and_assemble(system, U, UOld, F, matrix, ...)      src/vfvm_assembly.jl:520-534
 *     _solve_linear!(u, state, nlhistory, control, ...)       src/vfvm_linsolve.jl:6
 *     solve_step!(state, solution, oldsol, control, ...)      src/vfvm_solver.jl:13-23
 *     update_grid!(system)                                    src/vfvm_system.jl:607-631
 * A Julia shim adds methods for a device-backed SystemState and forwards to these entry points with `ccall`
 * (see INTEGRATION.md).  All entry points are `extern "C"`, take plain pointers and sizes, return an int
 * status (0 = ok, <0 = error, text via vfvm_last_error) and never call back into the host language.
 *
 * Conventions
 *   - node / edge / cell / dof indices are 0-based in this ABI (the Julia shim converts from 1-based);
 *     region numbers are labels and are passed through unchanged (1-based, as ExtendableGrids numbers them).
 *   - arrays are column-major exactly as the Julia arrays they mirror: coord is dim x N, cellnodes is
 *     (dim+1) x C, solution vectors are nspecies x N with dof = K*nspecies + ispec
 *     (src/vfvm_densesolution.jl:49).
 *   - caller owns every host buffer; the library owns every device buffer.
 *   - a handle is single-threaded (one SystemState is not re-entrant in the reference either,
 *     examples/Example440_ParallelState.jl:36-45); use one handle per state / per GPU.
 *   - `memspace` tells whether a vector argument points to host (VFVM_HOST) or device (VFVM_DEVICE) memory.
 */
#ifndef VFVM_B200_H
#define VFVM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vfvm_handle vfvm_handle;

/* ---- status codes (mapped by the shim onto src/vfvm_logging_exceptions.jl:5-29) ------------------- */
#define VFVM_OK 0
#define VFVM_ERR_ARG (-1)          /* bad argument                                                    */
#define VFVM_ERR_STATE (-2)        /* call order violated (e.g. assemble before build_pattern)         */
#define VFVM_ERR_CUDA (-3)         /* CUDA runtime failure                                             */
#define VFVM_ERR_NAN (-4)          /* "trying to assemble NaN" src/vfvm_assembly.jl:10-12 -> AssemblyError */
#define VFVM_ERR_LINSOLVE (-5)     /* Krylov breakdown / singular block -> LinearSolverError           */
#define VFVM_ERR_UNREGISTERED (-6) /* physics id not in the registered device library                  */
#define VFVM_ERR_COMM (-7)         /* NCCL failure                                                     */
#define VFVM_ERR_UNSUPPORTED (-8)  /* feature outside the hot-path scope         */

#define VFVM_HOST 0
#define VFVM_DEVICE 1

/* ---- coordinate systems: src/vfvm_xgrid.jl:6-47 selects the cellfactors! method ------------------- */
#define VFVM_CARTESIAN 0
#define VFVM_CYLINDRICAL 1 /* 1D: Polar1D, 2D: Cylindrical2D */
#define VFVM_SPHERICAL 2   /* 1D only: Spherical1D           */

/* ---- physics slots: fields of VoronoiFVM.Physics, src/vfvm_physics.jl:67-182 ----------------------- */
#define VFVM_SLOT_FLUX 0
#define VFVM_SLOT_REACTION 1
#define VFVM_SLOT_STORAGE 2
#define VFVM_SLOT_SOURCE 3
#define VFVM_SLOT_BREACTION 4
#define VFVM_SLOT_EDGEREACTION 5 /* edgereaction(f,u,edge,data), src/vfvm_assembly.jl:202-239 */
#define VFVM_SLOT_BSTORAGE 6     /* bstorage(f,u,bnode,data), src/vfvm_assembly.jl:409-439 */
#define VFVM_NUM_SLOTS 7

/*
 * ---- registered physics library ----------------------------------------------------------------------
 * Every callback the device can evaluate is one id + a parameter block of doubles (species indices are
 * stored as 0-based doubles).  n = number of species.  uK/uL = unknowns at edge.node[1]/edge.node[2].
 * An id that is not listed here is rejected with VFVM_ERR_UNREGISTERED -- there is no CPU fallback.
 */
#define VFVM_NONE 0

/* flux(f,u,edge,data) */
#define VFVM_FLUX_DIFFUSION 1   /* f_i = D_i (uK_i - uL_i)                 params: D[n]    (Example201:17-20, 301:17-20, 410:20-25) */
#define VFVM_FLUX_POWDIFF 2     /* f_i = D_i (uK_i^m - uL_i^m)             params: D[n], m (Example207:35-38, 106:49-52)            */
#define VFVM_FLUX_CROSSDIFF2 3  /* n=2: f_1 = e1 (u1K-u1L)(c+u2K+u2L), f_2 = e2 (u2K-u2L)(c+u1K+u1L)  params: e1,e2,c (Example110:43-50) */
#define VFVM_FLUX_SG_UNIPOLAR 4 /* Example160 classflux! :43-50            params: eps, iphi, ic                                     */
#define VFVM_FLUX_SEDAN 5       /* Example160 sedanflux! :68-77            params: eps, z, iphi, ic, eps_reg                         */
#define VFVM_FLUX_SG_BIPOLAR 6  /* Example161 flux! :134-150               params: lambda, mun, mup, zn, zp, En, Ep, iphin, iphip, ipsi */

#define VFVM_FLUX_MIXTURE 7     /* DevEx005_Mixture.jl:74-104: f = M(u)^{-1} (uK - uL) with the Maxwell-Stefan matrix M_ii = 1/DK_i + sum_{j!=i} au_j / DB_ij,
                                   M_ij = -au_i / DB_ij, au = (uK + uL)/2, solved in the callback by inplace_linsolve! (src/vfvm_functions.jl:98-168)
                                   params: DK[n], DB[n*n] row-major (diagonal ignored) */

/* reaction(f,u,node,data) */
#define VFVM_REACTION_POW 1       /* f_i = k_i u_i^{p_i}                    params: k[n], p[n]   (Example207:32-34)                  */
#define VFVM_REACTION_SINH 2      /* f_i = k_i (exp(u_i) - exp(-u_i))       params: k[n]         (Example105:55-58)                  */
#define VFVM_REACTION_AFFINE 3    /* f = R u + r0                           params: R[n*n] row-major, r0[n] (Example210:27-31, 160:60-66) */
#define VFVM_REACTION_BILINEAR2 4 /* n=2: f_1 = k u1 u2, f_2 = -k u1 u2     params: k            (Example110:38-42)                  */
#define VFVM_REACTION_REGION_AFFINE 6 /* f = R_r u + r0_r in cell region r  params: nreg, then per region R[n*n] row-major, r0[n] (Example221:54-66) */
#define VFVM_REACTION_BIPOLAR 5   /* Example161 reaction! :109-132          params: zn, zp, En, Ep, r0, iphin, iphip, ipsi, nreg, C[nreg] */

/* storage(f,u,node,data) */
#define VFVM_STORAGE_LINEAR 1  /* f_i = c_i u_i                             params: c[n]         (Example207:44-47, 160:52-58)       */
#define VFVM_STORAGE_POW 2     /* f_i = (eps_i + u_i)^(1/m_i)               params: eps[n], m[n] (Example107:52-55)                  */
#define VFVM_STORAGE_BIPOLAR 3 /* Example161 storage! :163-170              params: zn, zp, En, Ep, iphin, iphip, ipsi               */

/* source(f,node,data) */
#define VFVM_SOURCE_CONST 1    /* f_i = s_i                                 params: s[n]                                             */
#define VFVM_SOURCE_GAUSS 2    /* f_sp = exp(-a sum_d (x_d - c_d)^2)        params: sp, a, c[3]  (Example207:39-43, 210:39-44)       */
#define VFVM_SOURCE_XSINYEXPZ 3 /* f_sp = x sin(b y) exp(z)                 params: sp, b        (Example301:22-26)                  */
#define VFVM_SOURCE_STEP1D 4   /* f_sp = x <= x0 ? lo : hi                  params: sp, x0, lo, hi (Example105:45-52)                */
#define VFVM_SOURCE_AFFINE_X 5 /* f_i = a_i + b_i x                         params: a[n], b[n]   (Example110:51-55)                  */
#define VFVM_SOURCE_NODAL 6    /* f_i = table[i,K] uploaded by vfvm_set_nodal_source (host-evaluated, u-independent callback)         */

/* breaction(f,u,bnode,data): the part that is not a boundary_dirichlet!/neumann!/robin! call */
#define VFVM_BREACTION_LINEAR 1 /* if bnode.region == r: f = R u            params: r, R[n*n] row-major (Example215:33-42)           */

#define VFVM_BREACTION_CATALYSIS 2 /* Example115 breaction! :128-135: if bnode.region == r: f_A = S R_AC, f_B = S R_BC, f_C = -R_BC - R_AC with
                                     R_XC = kp_XC u_X (1 - u_C) - km_XC u_C   params: r, S, kpAC, kmAC, kpBC, kmBC, iA, iB, iC (0-based) */

#define VFVM_BREACTION_POW 3 /* if bnode.region == r: f_i = k_i u_i^{p_i}   params: r, k[n], p[n]   (Example226_BoundaryIntegral.jl:42-47: u^2) */

/* edgereaction(f,u,edge,data): a reaction term given per edge, assembled with the edge form factor into BOTH end nodes
 * (src/vfvm_assembly.jl:202-239) */
#define VFVM_EDGEREACTION_DIAMOND 1 /* f_i = c_i h^2 / (2 dim), h = meas(edge): a constant volume density per half diamond (DevEx002:83-87) params: c[n] */
#define VFVM_EDGEREACTION_JOULE 2   /* f_iT = -kappa (u_iphi,K - u_iphi,L)^2                (Example206:83-86)  params: kappa, iphi, iT (0-based) */

/* bstorage(f,u,bnode,data) */
#define VFVM_BSTORAGE_LINEAR 1 /* if bnode.region == r: f_i = c_i u_i      params: r, c[n]   (Example115:138-143, Example311:78-83) */

/* boundary condition entries = calls of the callback-level helpers src/vfvm_physics.jl:487-564 */
#define VFVM_BC_DIRICHLET 1 /* boundary_dirichlet!(y,u,bnode,ispec,ireg,val) :487-494 */
#define VFVM_BC_NEUMANN 2   /* boundary_neumann!(y,u,bnode,ispec,ireg,val)   :533     */
#define VFVM_BC_ROBIN 3     /* boundary_robin!(y,u,bnode,ispec,ireg,fac,val) :552     */

typedef struct vfvm_bc_entry {
    int32_t kind;     /* VFVM_BC_*                                                             */
    int32_t species;  /* 0-based                                                               */
    int32_t region;   /* boundary region label; 0 = all boundary regions (region=bnode.region) */
    int32_t has_ramp; /* value = ramp(bnode.time; dt=(t0,t1), du=(v0,v1)) src/vfvm_physics.jl:516-525 */
    double value;
    double factor; /* Robin alpha */
    double t0, t1, v0, v1;
} vfvm_bc_entry;

/* vectors of the device-resident SystemState (src/vfvm_state.jl:16-83) */
#define VFVM_VEC_SOLUTION 0
#define VFVM_VEC_OLDSOL 1
#define VFVM_VEC_RESIDUAL 2
#define VFVM_VEC_UPDATE 3

/* linear solver selection (replaces LinearSolve algorithms in SolverControl.method_linear,
 * src/vfvm_solvercontrol.jl:106; examples/Example207_NonlinearPoisson2D.jl:86, DevEx003_Solvers.jl:100-147) */
#define VFVM_KRYLOV_BICGSTAB 0
#define VFVM_KRYLOV_CG 1
#define VFVM_KRYLOV_GMRES 2
#define VFVM_PRECON_NONE 0
#define VFVM_PRECON_JACOBI 1      /* JacobiPreconBuilder: point diagonal                         */
#define VFVM_PRECON_BLOCKJACOBI 2 /* n x n node-block inverse (BlockPreconBuilder with node blocks) */
#define VFVM_PRECON_ILU0 3        /* ILUZeroPreconBuilder on the node-block pattern, natural order  */
#define VFVM_PRECON_ILU0_MC 4     /* same factorisation in multicolour elimination order (few levels) */
#define VFVM_PRECON_AMG 5         /* AMGPreconBuilder: aggregation AMG V-cycle on the node-block matrix (csrc/amg.cu)  */

/* ---- lifecycle -------------------------------------------------------------------------------------- */
int vfvm_create(int device, vfvm_handle** out);
void vfvm_destroy(vfvm_handle* h);
const char* vfvm_last_error(vfvm_handle* h);
int vfvm_abi_version(void);

/* ---- grid: the ExtendableGrids arrays read at src/vfvm_system.jl:691-700 ------------------------------ */
int vfvm_set_grid(vfvm_handle* h, int dim, int coordsys, int64_t nnodes, int64_t ncells, int64_t nbfaces,
                  const double* coord, const int32_t* cellnodes, const int32_t* cellregions,
                  const int32_t* bfacenodes, const int32_t* bfaceregions);
/* multi-GPU: the first n_owned local nodes are owned by this rank, the rest are halo copies */
int vfvm_set_owned_nodes(vfvm_handle* h, int64_t n_owned);

/* K1+K2: edge enumeration (grid[CellEdges], grid[EdgeNodes], src/vfvm_system.jl:702-703) and
 * update_grid_edgewise! (src/vfvm_system.jl:690-755) incl. cellfactors!/bfacefactors!
 * (src/vfvm_formfactors.jl:12-332) */
int vfvm_build_geometry(vfvm_handle* h);
int vfvm_num_edges(vfvm_handle* h, int64_t* nedges);
int vfvm_get_edgenodes(vfvm_handle* h, int32_t* edgenodes /* 2 x E */);
int vfvm_get_celledges(vfvm_handle* h, int32_t* celledges /* ne x C */);
/* nodefactors / edgefactors are nregions x N / nregions x E CSC matrices (src/vfvm_assemblydata.jl:39-55) */
int vfvm_num_factors(vfvm_handle* h, int64_t* n_nodefactors, int64_t* n_edgefactors);
int vfvm_get_nodefactors(vfvm_handle* h, int64_t* colptr, int32_t* region, double* fac);
int vfvm_get_edgefactors(vfvm_handle* h, int64_t* colptr, int32_t* region, double* fac);
int vfvm_get_bfacefactors(vfvm_handle* h, double* bfacenodefactors /* dim x NB */);

/* ---- system: enable_species! (src/vfvm_system.jl:433-480), physics!, legacy BC tables (:854-933) ------ */
int vfvm_set_system(vfvm_handle* h, int nspecies, const uint8_t* region_species /* n x nregions or NULL = all */);
/* enable_boundary_species!(sys, ispec, bregions) (src/vfvm_system.jl:492-515): species that live on boundary regions only; n x nbregions,
 * column-major, NULL = none.  Call after vfvm_set_system and before vfvm_build_pattern. */
int vfvm_set_boundary_species(vfvm_handle* h, int nbregions, const uint8_t* bregion_species);
int vfvm_set_physics(vfvm_handle* h, int slot, int physics_id, const double* params, int nparams);
int vfvm_set_nodal_source(vfvm_handle* h, const double* table /* n x N, host */);
int vfvm_set_legacy_bc(vfvm_handle* h, int nbregions, const double* boundary_factors /* n x nbregions */,
                       const double* boundary_values);
int vfvm_set_bc_entries(vfvm_handle* h, int nentries, const vfvm_bc_entry* entries);

/* K3: sparsity pattern + maps (what rawupdateindex!/flush! of ExtendableSparse build implicitly,
 * src/vfvm_assembly.jl:26, src/vfvm_solver.jl:242).  Scalar pattern in dof numbering. */
int vfvm_build_pattern(vfvm_handle* h);
int vfvm_pattern_size(vfvm_handle* h, int64_t* nrows, int64_t* nnz);
int vfvm_get_pattern_csr(vfvm_handle* h, int64_t* rowptr, int64_t* colidx);
int vfvm_get_pattern_csc(vfvm_handle* h, int64_t* colptr, int64_t* rowval);

/* ---- state vectors ---------------------------------------------------------------------------------- */
int vfvm_set_vector(vfvm_handle* h, int which, const double* src, int memspace);
int vfvm_get_vector(vfvm_handle* h, int which, double* dst, int memspace);
int vfvm_copy_vector(vfvm_handle* h, int dst_which, int src_which);
/* _initialize_dirichlet! on the resident solution (src/vfvm_system.jl:947-1003) */
int vfvm_init_dirichlet(vfvm_handle* h, double time, double lambda);

/* ---- K4-K6: eval_and_assemble (src/vfvm_assembly.jl:520-643) ---------------------------------------- */
/* assembles residual + Jacobian at the resident SOLUTION / OLDSOL vectors; tstep = Inf for stationary */
int vfvm_assemble(vfvm_handle* h, double time, double tstep, double lambda);
/* same, but only enqueued on the handle's stream (SURVEY 8b "unless flags & VFVM_ASYNC"): a NaN is reported by the next
 * vfvm_sync; the following vfvm_linsolve / vfvm_newton_update run on the same stream and need no host round trip */
int vfvm_assemble_async(vfvm_handle* h, double time, double tstep, double lambda);
int vfvm_sync(vfvm_handle* h);
/* convenience = evaluate_residual_and_jacobian! (src/vfvm_solver.jl:224-243): upload U (and UOld, may be
 * NULL = U), assemble, download residual; the Jacobian stays on the device for the linear solve */
int vfvm_eval_res_jac(vfvm_handle* h, const double* U, const double* UOld, double* F, int memspace,
                      double time, double tstep, double lambda);
int vfvm_get_nzval_csr(vfvm_handle* h, double* nzval, int memspace);
int vfvm_get_nzval_csc(vfvm_handle* h, double* nzval, int memspace);
/* scalar CSR rows (pattern + values) of the owned nodes [node0, node1) only: the rows SparseMatrixCSC(flush!(matrix)) holds for these
 * nodes (src/vfvm_solver.jl:242), without moving the whole Jacobian to the host.  Call with colidx == NULL to get *nnz first;
 * rowptr has (node1 - node0) * nspecies + 1 offsets starting at 0, colidx are local dof numbers. */
int vfvm_get_rows_csr(vfvm_handle* h, int64_t node0, int64_t node1, int64_t* nnz, int64_t* rowptr, int64_t* colidx, double* nzval);

/* ---- K8-K10: _solve_linear! (src/vfvm_linsolve.jl:6-61): solve A * UPDATE = RESIDUAL --------------- */
int vfvm_linsolve_setup(vfvm_handle* h, int krylov, int precon, int gmres_restart);
int vfvm_linsolve(vfvm_handle* h, double abstol, double reltol, int maxiters, int reuse_precs, int* iters,
                  double* resnorm);
/* did the last vfvm_linsolve reach its tolerance (1) or stop at maxiters (0)?  rhs_norm = ||RESIDUAL||_2 it was measured against.
 * Hitting maxiters is not an error of vfvm_linsolve (Krylov.jl under LinearSolve behaves the same, the Newton loop judges the update);
 * the stand-in for the reference's default direct solver (UMFPACK, src/vfvm_solver.jl:34-41) turns it into a LinearSolverError. */
int vfvm_linsolve_status(vfvm_handle* h, int* converged, double* rhs_norm);
/* y = A x for parity tests of the SpMV kernel; x, y in the given memspace, length = nrows */
int vfvm_spmv(vfvm_handle* h, const double* x, double* y, int memspace);

/* options of VFVM_PRECON_AMG: opts = {omega, alpha, theta, sweeps, coarse_sweeps, wdepth}, NaN (or a shorter array) keeps a value;
 * the keyword arguments of AMGPreconBuilder on the host side */
int vfvm_amg_set_options(vfvm_handle* h, const double* opts, int nopts);

/* ---- K11: Newton update + norms (src/vfvm_solver.jl:116-132) ----------------------------------------- */
/* SOLUTION -= damp * UPDATE; returns ||UPDATE||_inf and ||SOLUTION_new||_1 */
int vfvm_newton_update(vfvm_handle* h, double damp, double* update_norm_inf, double* solution_norm1);
int vfvm_vector_norms(vfvm_handle* h, int which, double* norm_inf, double* norm1);
/* ||a - b||_inf: SolverControl.delta (src/vfvm_solvercontrol.jl:257) */
int vfvm_vector_diffnorm(vfvm_handle* h, int which_a, int which_b, double* norm_inf);

/* ---- post-processing integrals of a resident vector (src/vfvm_postprocess.jl:18-67, :109-146) -------------
 * out: n x ncellregions (column-major) on the host.  vfvm_integrate: slot = VFVM_SLOT_REACTION or VFVM_SLOT_STORAGE selects the
 * evaluator of a registered node function, id = VFVM_NONE integrates the vector itself (integrate(system, U)).
 * vfvm_edgeintegrate: id = a registered flux id, or -1 = the W^{1,p} seminorm integrand dim ((u_K - u_L) / h)^p, p = params[0]
 * (w1pseminorm, :300-312), -2 = the edge average (u_K + u_L) / 2 (test/test120_norms.jl:35-38).  Collective with several ranks. */
int vfvm_integrate(vfvm_handle* h, int slot, int id, const double* params, int np, int which, double* out);
int vfvm_edgeintegrate(vfvm_handle* h, int id, const double* params, int np, int which, double* out);
/* integrate(system, F, U; boundary = true) (src/vfvm_postprocess.jl:29-46): out is n x nbfaceregions (host); slot selects the evaluator of the
 * registered function (VFVM_SLOT_BREACTION, VFVM_SLOT_REACTION, VFVM_SLOT_STORAGE, VFVM_SLOT_BSTORAGE), id = VFVM_NONE integrates the vector itself;
 * a species contributes at the boundary nodes where it is defined (isnodespecies).  Collective with several ranks. */
int vfvm_integrate_boundary(vfvm_handle* h, int slot, int id, const double* params, int np, int which, double* out);
/* mass_matrix(state), src/vfvm_diffeq_interface.jl:60-101: Jacobian of the registered storage at U = 0 times the node factors;
 * out (host): one n x n block per owned node, out[(K*n + i)*n + j] = M[(K,i),(K,j)].  eval_rhs! / eval_jacobian! of the ODE
 * interface (:27-52) are vfvm_eval_res_jac with tstep = Inf and a sign flip on the host side. */
int vfvm_mass_matrix(vfvm_handle* h, double* out);
/* flux callback of every edge of a resident vector, without form factor: out (host) n x E, out[e*n + i] = flux(u_K, u_L)_i with
 * K = edge.node[1], L = edge.node[2] -- the edge loop of nodeflux (src/vfvm_postprocess.jl:191-207); the accumulation over the
 * nodes with the Voronoi face centres is host-side post-processing (voronoifvm.jl_b200/postprocess.py:nodeflux) */
int vfvm_edgeflux(vfvm_handle* h, int id, const double* params, int np, int which, double* out);

/* ---- multi-GPU (one process per GPU; the host shares the NCCL id through its own rendezvous) ------- */
int vfvm_comm_unique_id(char id_out[128]);
int vfvm_comm_init(vfvm_handle* h, int rank, int nranks, const char id[128]);
/* neighbour r sends us halo nodes [recv_ptr[r], recv_ptr[r+1]) (offsets into the halo range) and receives our
 * owned nodes send_idx[send_ptr[r] .. send_ptr[r+1]) */
int vfvm_set_halo(vfvm_handle* h, int nneighbors, const int32_t* neighbor_ranks, const int64_t* send_ptr,
                  const int32_t* send_idx, const int64_t* recv_ptr);
int vfvm_halo_exchange(vfvm_handle* h, int which);
/* Peer-memory transport (CUDA IPC mailboxes over NVLink, csrc/peer.cuh): after vfvm_set_halo and vfvm_set_system every rank
 * exports its mailbox (64-byte IPC handle), the host gathers the handles of all ranks (rank order) and every rank connects.
 * From then on the halo exchange is part of the SpMV kernel and the Krylov reductions are part of their finalize kernel;
 * NCCL stays the transport if vfvm_peer_connect fails (returns VFVM_ERR_COMM) or VFVM_NO_PEER is set. */
int vfvm_peer_export(vfvm_handle* h, char handle_out[64]);
int vfvm_peer_connect(vfvm_handle* h, const char* handles /* nranks x 64 bytes */);
int vfvm_peer_active(vfvm_handle* h);

/* ---- instrumentation --------------------------------------------------------------------------------- */
#define VFVM_TIME_ASSEMBLE 0       /* last vfvm_assemble, ms, CUDA events on the handle's stream (tasm)      */
#define VFVM_TIME_LINSOLVE_SETUP 1 /* tlinsolve_setup src/vfvm_history.jl:17-30                              */
#define VFVM_TIME_LINSOLVE_SOLVE 2 /* tlinsolve_solve                                                        */
#define VFVM_TIME_EDGE_KERNEL 3    /* last row-tile (edge+node) kernel alone, ms                             */
#define VFVM_NUM_TIMES 8
int vfvm_timings(vfvm_handle* h, double* ms_out /* VFVM_NUM_TIMES */);
int vfvm_launch_count(vfvm_handle* h, int64_t* nlaunches); /* kernels launched by this handle so far */
int vfvm_stream(vfvm_handle* h, void** cuda_stream);       /* the cudaStream_t all work is enqueued on */
int vfvm_device_bytes(vfvm_handle* h, int64_t* bytes);
/* stored Jacobian planes: species couplings kept per off-diagonal block (flux mask) / per diagonal block */
int vfvm_plane_counts(vfvm_handle* h, int* off_planes, int* diag_planes);
/* off-diagonal blocks of the owned rows (= 2 x edges, summed over ranks) and blocks stored incl. SELL-32 padding */
int vfvm_block_counts(vfvm_handle* h, int64_t* nblocks_off, int64_t* nblocks_stored);

/* parity probe for test/test010_bernoulli.jl: evaluates the device fbernoulli_pm (src/vfvm_functions.jl:78-90) and its
 * dual-number derivative at n host points: bp = B(x), bm = B(-x), dbp = B'(x) */
int vfvm_probe_bernoulli(vfvm_handle* h, int n, const double* x, double* bp, double* bm, double* dbp);
/* parity probe for test/test040_inplacelu.jl: the device twins of inplace_linsolve!(A, b) (non-pivoting Doolittle, src/vfvm_functions.jl:98-155)
 * and inplace_linsolve!(A, b, ipiv) (pivoting LU, :165-168) on `nsys` systems of size n x n (n <= 10; A row-major, nsys x n x n; b, x: nsys x n) */
int vfvm_probe_inplace_linsolve(vfvm_handle* h, int n, int nsys, int pivoting, const double* A, const double* b, double* x);

#ifdef __cplusplus
}
#endif
#endif /* VFVM_B200_H */
