/* cedr_b200 -- C ABI of the B200-native CEDR property-preservation hot path.
 *
 * This is the drop-in boundary for COMPOSE's `cedr::CDR` interface as realised
 * by QLT (cedr/cedr_qlt.{hpp,cpp}) and CAAS (cedr/cedr_caas.{hpp,cpp}). Each
 * entry point names the reference interface it replaces; `file:line` are
 * relative to the reference's cedr/ directory. The C++ mirror of the
 * reference's classes (include/cedr_b200.hpp) and the Python front end
 * (compose_b200/__init__.py) are thin layers over exactly these symbols.
 *
 * Conventions
 *   - every function returns 0 on success and a nonzero code on failure;
 *     cedr_b200_last_error() then returns a message formatted like the
 *     reference's cedr_throw_if (cedr_util.hpp:70-77). The C++ mirror re-throws
 *     it as std::logic_error.
 *   - all `double*` data pointers are DEVICE pointers unless the name says
 *     `host`. Sizes are counts of doubles, as in the reference.
 *   - work is enqueued on the CDR's stream (cedr_b200_set_stream, default: the
 *     legacy default stream). cedr_b200_run() is asynchronous; the reference's
 *     run() is synchronous, so the C++ mirror follows it with
 *     cedr_b200_synchronize().
 *   - there is no CPU fallback: if no CUDA device is usable, creation fails.
 */
#ifndef CEDR_B200_H
#define CEDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cedr.hpp:29-39, ProblemType */
enum {
  CEDR_B200_CONSERVE = 1,
  CEDR_B200_SHAPEPRESERVE = 2,
  CEDR_B200_CONSISTENT = 4,
  CEDR_B200_NONNEGATIVE = 8
};

/* How CAAS forms its four per-tracer global sums (cedr_caas.cpp:129-209). */
enum {
  /* Pairwise in the order of a recursive-bisection tree over the cells: what
   * the reference computes when its UserAllReducer is backed by
   * BfbTreeAllReducer (cedr_bfb_tree_allreduce.cpp:86-124). Deterministic and
   * independent of the GPU decomposition. Default. */
  CEDR_B200_CAAS_SUM_TREE = 0,
  /* Sequential over cells 0..n-1: what the reference's own reduce_locally
   * computes on a host backend (team size 1, cedr_kokkos.hpp:118). One rank; a
   * compatibility mode (one thread per tracer), bit-identical to the reference's
   * default CAAS. */
  CEDR_B200_CAAS_SUM_SEQUENTIAL = 1,
  /* By the caller's UserAllReducer (cedr_caas.hpp:27-49, cedr_caas.cpp:140-168, 262-266):
   * CAAS hands it nlclcells / n_accum_in_place partial sums per field and uses the 4 nt
   * sums it returns. The reducer owns the cross-rank reduction, so this rank's cells may be
   * any set (cell0 / ncells_global are not used). See cedr_b200_caas_set_user_reducer. */
  CEDR_B200_CAAS_SUM_USER = 2
};

typedef struct cedr_b200_cdr cedr_b200_cdr;

/* Message of the last failure on this thread. */
const char* cedr_b200_last_error(void);
/* Library/ABI version, and whether a usable CUDA device is present (0/1). */
int cedr_b200_version(void);
int cedr_b200_device_available(void);

/* ---- construction ------------------------------------------------------ */

/* QLT<ES>::QLT(p, ncells, tree, options), cedr_qlt.cpp:229-236.
 * The caller's tree::Node graph (cedr_tree_caller.hpp:12-24) is passed
 * flattened: node i has kids[2*i], kids[2*i+1] (both -1 for a leaf);
 * cellidx[i] is the global cell of leaf i; node_rank[i] the owning rank of leaf
 * i (NULL: all 0). `rank`/`nranks` identify this process (Parallel::rank/size,
 * cedr_mpi.hpp:17-27). Fails if this rank owns no cell (cedr_qlt.cpp:235). */
int cedr_b200_qlt_create(cedr_b200_cdr** cdr, int ncells, int nnodes, int root,
                         const int* kids, const int64_t* cellidx,
                         const int* node_rank,
                         int prefer_numerical_mass_conservation_to_numerical_bounds,
                         int rank, int nranks);

/* Same, over tree::make_tree_over_1d_mesh(p, ncells, imbalanced)
 * (cedr_tree_caller.hpp:26-29, cedr_tree.cpp:350-353) with the contiguous
 * cell->rank map of oned::Mesh (cedr_tree.cpp:366-369). */
int cedr_b200_qlt_create_1d(cedr_b200_cdr** cdr, int ncells, int imbalanced,
                            int prefer_numerical_mass_conservation_to_numerical_bounds,
                            int rank, int nranks);

/* CAAS<ES>::CAAS(p, nlclcells, UserAllReducer::Ptr), cedr_caas.cpp:37-48.
 * `sum_mode` is one of CEDR_B200_CAAS_SUM_*. `cell0`/`ncells_global` place this
 * rank's cells in the global cell order (tree-ordered sums need it); one rank:
 * cell0 = 0, ncells_global = nlclcells. */
int cedr_b200_caas_create(cedr_b200_cdr** cdr, int nlclcells, int sum_mode,
                          int64_t cell0, int64_t ncells_global, int rank,
                          int nranks);

/* CAAS::UserAllReducer::operator() (cedr_caas.hpp:31-39): an MPI_Allreduce-like call.
 * send is (nlocal fastest, nfld), recv is (nfld), both DEVICE pointers as on the
 * reference's GPU builds; the operation is always a sum (the reference passes MPI_SUM,
 * cedr_caas.cpp:262-266). `stream` is the CDR's CUDA stream: it has been synchronised
 * before the call, and whatever the reducer enqueues must be complete or ordered on
 * `stream` when it returns. The implementation may modify send. Return 0 on success. */
typedef int (*cedr_b200_user_reducer_fn)(void* ctx, double* send, double* recv, int nlocal,
                                         int nfld, void* stream);
/* The `r` argument of CAAS<ES>::CAAS (cedr_caas.cpp:37-48); `n_accum_in_place` is
 * UserAllReducer::n_accum_in_place() (cedr_caas.hpp:41-48) and must divide nlclcells (the
 * reference silently drops the remainder cells, cedr_caas.cpp:139). Only for a CAAS created
 * with CEDR_B200_CAAS_SUM_USER; call before end_tracer_declarations. */
int cedr_b200_caas_set_user_reducer(cedr_b200_cdr* cdr, cedr_b200_user_reducer_fn fn,
                                    void* ctx, int n_accum_in_place);

/* ~CDR */
int cedr_b200_destroy(cedr_b200_cdr* cdr);

/* ---- CDR interface, cedr_cdr.hpp:16-112 -------------------------------- */

/* CDR::declare_tracer (cedr_qlt.cpp:276-285, cedr_caas.cpp:50-58) */
int cedr_b200_declare_tracer(cedr_b200_cdr* cdr, int problem_type, int rhomidx);
/* CDR::end_tracer_declarations (cedr_qlt.cpp:287-291, cedr_caas.cpp:60-73) */
int cedr_b200_end_tracer_declarations(cedr_b200_cdr* cdr);
/* CDR::get_buffers_sizes (cedr_qlt.cpp:293-298, cedr_caas.cpp:75-90) */
int cedr_b200_get_buffers_sizes(cedr_b200_cdr* cdr, size_t* buf1, size_t* buf2);
/* CDR::set_buffers (cedr_qlt.cpp:300-305, cedr_caas.cpp:92-99): device memory of
 * at least the sizes above; must outlive the CDR. */
int cedr_b200_set_buffers(cedr_b200_cdr* cdr, double* buf1, double* buf2);
/* CDR::finish_setup (cedr_qlt.cpp:307-313, cedr_caas.cpp:101-116) */
int cedr_b200_finish_setup(cedr_b200_cdr* cdr);
/* CDR::get_problem_type: the canonical type (cedr_qlt.cpp:315-320) */
int cedr_b200_get_problem_type(const cedr_b200_cdr* cdr, int tracer_idx, int* type);
/* CDR::get_num_tracers */
int cedr_b200_get_num_tracers(const cedr_b200_cdr* cdr, int* ntracers);
/* CDR::run (cedr_qlt.cpp:618-640, cedr_caas.cpp:258-270). Asynchronous. */
int cedr_b200_run(cedr_b200_cdr* cdr);
/* CDR::print */
int cedr_b200_print(const cedr_b200_cdr* cdr, char* buf, size_t bufsize);

/* QLT::nlclcells / get_owned_glblcells / gci2lci, cedr_qlt.cpp:241-274 */
int cedr_b200_nlclcells(const cedr_b200_cdr* cdr, int* n);
int cedr_b200_get_owned_glblcells(const cedr_b200_cdr* cdr, int64_t* gcis_host);
int cedr_b200_gci2lci(const cedr_b200_cdr* cdr, int64_t gci, int* lci);

/* ---- DeviceOp, cedr_cdr.hpp:67-102 ------------------------------------- */

/* Trivially copyable view of the CDR's device buffers. Caller kernels copy it
 * by value and call the inline set_rhom / set_Qm / get_Qm of
 * include/cedr_b200_device_op.h on it, exactly as they copy the reference's
 * concrete DeviceOp into Kokkos lambdas (cedr_test_randomized_inl.hpp:27-58).
 * Layout: SoA, cell fastest: word `row` of local cell `lci` is data[row*ld + lci].
 * Row 0 of `in` is rhom; tracer t owns rows trcr_row[t]..: Qm_min (or q_min),
 * Qm, Qm_max (or q_max), Qm_prev as its problem type needs
 * (cedr_qlt_inl.hpp:21-58). QLT results land in out[t*ld + lci]; CAAS works in
 * place (out aliases the Qm rows of `in`), like the reference. */
typedef struct cedr_b200_device_op {
  double* in;
  double* out;
  int64_t ld;
  const int* trcr_row;   /* device, ntracers entries */
  const int* trcr_prob;  /* device, canonical problem types */
  int ntracers;
  int nlclcells;
  int is_caas;
  int reserved;
} cedr_b200_device_op;

int cedr_b200_get_device_op(cedr_b200_cdr* cdr, cedr_b200_device_op* op);

/* Bulk forms of DeviceOp::set_rhom / set_Qm / get_Qm over SoA caller arrays
 * (SURVEY.md section 8f-1): a[t*lda + lci] for t in [t0, t0+nt). Device
 * pointers; qm_prev may be NULL when none of the tracers conserves. */
int cedr_b200_set_rhom_bulk(cedr_b200_cdr* cdr, const double* rhom);
int cedr_b200_set_Qm_bulk(cedr_b200_cdr* cdr, int t0, int nt, int64_t lda,
                          const double* qm, const double* qm_min,
                          const double* qm_max, const double* qm_prev);
int cedr_b200_get_Qm_bulk(cedr_b200_cdr* cdr, int t0, int nt, int64_t lda,
                          double* qm);

/* ---- streams, multi-GPU exchange --------------------------------------- */

/* Zero-copy DeviceOp::set_Qm / get_Qm (cedr_qlt_inl.hpp:21-66, cedr_caas_inl.hpp:21-42;
 * the caller kernels of cedr_test_randomized_inl.hpp:32-58): instead of scattering its SoA
 * arrays into the CDR's buffer and gathering the results back, the caller binds them and
 * run() reads them in place -- a[t*lda + lci], tracer-major, local cell index fastest --
 * and writes QLT's results to qm_out (CAAS works in place on qm, as the reference does on
 * its buffer). All tracers must be shape-preserving (consistent-only tracers store scaled
 * bounds, nonnegative ones have none: use set_Qm for those). Arrays must stay valid and,
 * where the fast kernels apply, be 16-byte aligned with an even lda. qm == NULL unbinds.
 * rhom is still set through set_rhom. */
int cedr_b200_bind_arrays(cedr_b200_cdr* cdr, int64_t lda, const double* qm_min, double* qm,
                          const double* qm_max, const double* qm_prev, double* qm_out);

/* All work of this CDR is enqueued on `cuda_stream` (a cudaStream_t). */
int cedr_b200_set_stream(cedr_b200_cdr* cdr, void* cuda_stream);
int cedr_b200_synchronize(cedr_b200_cdr* cdr);

/* The one exchange step of a multi-GPU run() (replaces the per-level MPI
 * messages of cedr_qlt.cpp:327-337, 432-439, 478-488, 606-613 and CAAS's
 * MPI_Allreduce, cedr_caas.cpp:203-209): an all-gather, enqueued on `stream`,
 * of `count` doubles from every rank's `send` into `recv` (rank-major). The
 * plug-in point mirrors CAAS::UserAllReducer (cedr_caas.hpp:27-49): the host
 * side supplies NCCL (torch.distributed in Python, ncclAllGather in C++).
 * Must be set on every rank when nranks > 1. */
typedef int (*cedr_b200_allgather_fn)(void* ctx, const double* send, double* recv,
                                      size_t count, void* cuda_stream);
int cedr_b200_set_allgather(cedr_b200_cdr* cdr, cedr_b200_allgather_fn fn, void* ctx);

/* Multi-rank partition: rank r owns the cells the tree assigns to it
 * (tree::Node::rank, cedr_tree_caller.hpp:14); they must form whole tier-0 blocks of the
 * plan (a subtree partition, e.g. the contiguous map of cedr_tree.cpp:368-369 on a
 * bisection tree). Each rank up-sweeps its own blocks, the block roots (record + rhom)
 * are all-gathered -- `count` doubles per rank, the same on every rank -- and every rank
 * then sweeps the replicated tiers above in the fixed tree order, so results are
 * bit-identical to a one-rank run (the reference gets the same property from
 * cedr_bfb_tree_allreduce.cpp:115-124). get_exchange_count is valid after
 * end_tracer_declarations. set_exchange_buffers (before the first run) hands in caller
 * device buffers of `count` and `nranks*count` doubles, e.g. registered NCCL or
 * torch tensors; otherwise they are allocated internally. */
int cedr_b200_get_exchange_count(const cedr_b200_cdr* cdr, size_t* count);
int cedr_b200_set_exchange_buffers(cedr_b200_cdr* cdr, double* send, double* recv);
int cedr_b200_get_exchange_buffers(const cedr_b200_cdr* cdr, double** send, double** recv);
/* run() split at the exchange, for callers that drive the collective themselves and for
 * emulating P ranks on one device: phase 0 enqueues everything up to the packed message
 * in `send`; the caller fills `recv` with every rank's message (rank-major); phase 1
 * enqueues the rest. On one rank phase 0 is the whole run() and phase 1 does nothing. */
int cedr_b200_run_phase(cedr_b200_cdr* cdr, int phase);
/* Peer-to-peer exchange (one process per GPU, NVLink / NVSwitch): instead of handing the
 * message to an all-gather, the pack kernel stores it straight into every rank's receive
 * buffer over peer-mapped memory, and a one-warp kernel publishes and awaits an epoch flag
 * with system-scope release/acquire -- no library collective and no host involvement in
 * run(). Set-up, after finish_setup on every rank: exchange the 64-byte handles
 * (cedr_b200_p2p_get_handle; CUDA IPC) by any means, hand each peer's to
 * cedr_b200_p2p_set_peer, then cedr_b200_p2p_enable(cdr, 1). Receive buffers are
 * double-buffered by epoch parity, so a rank may start its next run() while a peer still
 * reads the previous message. Up to 16 ranks on one node. */
int cedr_b200_p2p_get_handle(cedr_b200_cdr* cdr, void* handle64);
int cedr_b200_p2p_set_peer(cedr_b200_cdr* cdr, int peer_rank, const void* handle64);
int cedr_b200_p2p_enable(cedr_b200_cdr* cdr, int on);
/* Host-only view of the partition of make_tree_over_1d_mesh(ncells, imbalanced) with the
 * contiguous cell->rank map, for `rank` of `nranks`: number of owned cells and blocks,
 * the padded block count of the exchange, and for each owned block (up to `cap`) its
 * global block index, first leaf (global DFS index) and leaf count. */
int cedr_b200_partition_probe(int ncells, int imbalanced, int max_block_leaves, int rank,
                              int nranks, int cap, int* nlclcells, int* nown, int* nown_max,
                              int* nblocks_global, int* gidx_host, int* leaf0_host,
                              int* nl_host);

/* ---- BfbTreeAllReducer, cedr_bfb_tree_allreduce.hpp:15-55 ----------------- */

/* BfbTreeAllReducer<ES>(p, tree, nleaf, nfield): an all-reduce (sum) of `nfield` scalars
 * per leaf whose result is bit-for-bit independent of the rank decomposition, because it
 * adds in the order of the tree (every node: d = 0; d += kid0; d += kid1,
 * cedr_bfb_tree_allreduce.cpp:115-124). Device-resident here (the reference stages
 * through the host on GPUs, :55-76): the same block plan and sweep kernels as QLT's
 * up-sweep, one all-gather of block roots when nranks > 1. The tree is given flattened as
 * for cedr_b200_qlt_create; kids == NULL builds make_tree_over_1d_mesh(nleaf).
 * max_block_leaves <= 0 keeps the default. The handle is destroyed with
 * cedr_b200_destroy and takes cedr_b200_set_stream / set_allgather / exchange buffers
 * (before the first allreduce) like a CDR. */
int cedr_b200_bfb_create(cedr_b200_cdr** reducer, int nleaf, int nnodes, int root,
                         const int* kids, const int64_t* cellidx, const int* node_rank,
                         int nfield, int max_block_leaves, int rank, int nranks);
/* BfbTreeAllReducer::allreduce(send, recv, transpose), :78-159. Device pointers; send is
 * (nfield fastest, nlocal) or, if transpose, (nlocal fastest, nfield), leaves in local
 * leaf order; recv gets nfield sums (send and recv may alias). Asynchronous on the
 * handle's stream. phase = -1 does everything; 0 / 1 split at the exchange as
 * cedr_b200_run_phase does. */
int cedr_b200_bfb_allreduce(cedr_b200_cdr* reducer, const double* send, double* recv,
                            int transpose, int phase);

/* ---- element-local solvers, cedr_local.hpp:23-58 ------------------------ */

/* cedr::local::solve_1eq_bc_qp / caas / solve_1eq_nonneg / solve_1eq_bc_qp_2d
 * (cedr_local_inl.hpp:68-330) for a batch of `nprob` independent elements of `n` <= 16
 * values each, one thread per element with the element held in registers
 * (include/cedr_b200_local.hpp). Entry i of element p of every array is at
 * [i*stride_var + p*stride_prob]: stride_var = nprob, stride_prob = 1 is the coalesced
 * SoA layout; stride_var = 1, stride_prob = n is the reference's contiguous n-vectors.
 * b is [nprob]; info [nprob] receives the reference's return codes (0 for CAAS). w may be
 * NULL (all ones) where the method does not use it; xlo / xhi are ignored by the
 * nonnegative methods. max_its <= 0 means the reference's default (100). */
enum {
  CEDR_B200_LOCAL_QP = 0,           /* solve_1eq_bc_qp */
  CEDR_B200_LOCAL_CAAS = 1,         /* caas */
  CEDR_B200_LOCAL_NONNEG_LS = 2,    /* solve_1eq_nonneg, Method::least_squares */
  CEDR_B200_LOCAL_NONNEG_CAAS = 3,  /* solve_1eq_nonneg, Method::caas */
  CEDR_B200_LOCAL_QP_2D = 4         /* solve_1eq_bc_qp_2d (n = 2) */
};
int cedr_b200_local_solve(int method, int nprob, int n, const double* w, const double* a,
                          const double* b, const double* xlo, const double* xhi,
                          const double* y, double* x, int* info, int64_t stride_var,
                          int64_t stride_prob, int max_its, int clip, void* stream);

/* ---- 1-D transport harness, cedr_test_1d_transport.cpp:136-255 ------------ */

/* Problem1D::cycle(nsteps, y0, yf, cdr) on the device for tracer 0 of `cdr` (a QLT over
 * the 1-D mesh tree of ncells cells, or a CAAS of ncells cells; uniform mesh): every step is
 * the periodic cubic interpolation at the departure points fused with the caller side of
 * run_cdr (bounds over the domain of dependence, set_Qm), CDR::run, and get_Qm / area --
 * three launches. With use_graph = 1 the steps are captured once in a CUDA graph and
 * replayed (the many-tiny-calls pattern of BASELINE.json's config 5); with use_graph = 2
 * the whole cycle is ONE launch: a persistent CTA runs the same three phases per step with
 * block barriers in place of kernel boundaries (problems of a single block of at most 256
 * cells with one tracer, e.g. the reference's 111-cell test; an error otherwise). Results
 * are identical in all three forms. y0_host / yf_host hold
 * ncells + 1 values (the last is the periodic image of the first). ms_per_step (may be
 * NULL) receives the device time per step. rhom is set to the cell areas. */
int cedr_b200_transport1d_cycle(cedr_b200_cdr* cdr, int nsteps, const double* y0_host,
                                double* yf_host, int use_graph, float* ms_per_step);

/* ---- introspection for tests / benches --------------------------------- */

/* run() as a replayed CUDA graph: its launches are captured once (per exchange-buffer parity)
 * and replayed, so a run() of a dozen small kernels costs one launch on the host and no
 * gaps between kernels on the device. mode -1 (default): where it pays, i.e. multi-rank
 * runs over the peer-to-peer exchange and one-rank runs of at most 3.2e7 cell x tracer
 * updates (ne30 x 72 x 40 is 1.6e7); 0: never; 1: whenever run() is pure stream work (no
 * profiling, no all-gather hook or UserAllReducer, not the ring kernel, not the one-launch
 * tiny-problem path). The first run() after setup or after a change of bindings stays
 * plain. Results are identical either way. (No reference counterpart: the reference's
 * run() is host code, cedr_qlt.cpp:618-640.) */
int cedr_b200_set_graph (cedr_b200_cdr* c, int mode);
int cedr_b200_uses_graph (const cedr_b200_cdr* c, int* on);

/* Number of kernels launched by the last run() on this CDR. */
int cedr_b200_last_run_launches(const cedr_b200_cdr* cdr, int* n);
/* Tier-0 kernel selection. Blocks shaped like a recursive bisection with 513..1024
 * leaves (every block of the cubed-sphere configs) run the fast kernels
 * (TMA-staged, register micro-subtrees); any other tree runs the generic
 * shared-memory kernels. Both produce identical bits. set_fast_path(0) before
 * finish_setup forces the generic kernels (A/B testing); uses_fast_path reports
 * the choice after finish_setup. */
int cedr_b200_set_fast_path(cedr_b200_cdr* cdr, int on);
int cedr_b200_uses_fast_path(const cedr_b200_cdr* cdr, int* on);
/* Opt-in (set_ring(1) before finish_setup) for the shape-preserving QLT classes and CAAS
 * where the fast shapes apply on one rank: run() as ONE persistent cooperative kernel per
 * problem class (ring_kernels.cuh) in place of QLT::run's 2*nlev+1 launches
 * (cedr_qlt.cpp:618-640) and CAAS::run's three (cedr_caas.cpp:258-270). Every CTA owns a
 * fixed piece of the leaves; the down-sweep re-reads them from L2 a few tracers behind the
 * up-sweep, so the leaves cross HBM once. Bit-identical to the multi-launch path, which
 * stays the default because it is faster today. uses_ring reports
 * the choice; ring_info returns {grid, sub-root depth, depth-7 nodes per piece, tracers
 * per unit, 100 x UP slots + DOWN slots, L groups, S warps, dynamic smem bytes}. */
int cedr_b200_set_ring(cedr_b200_cdr* cdr, int on);
/* Opt-in: CAAS::run (cedr_caas.cpp:258-270: reduce_locally, reduce_globally, finish_locally)
 * as ONE kernel of thread-block clusters (cluster_caas.cuh): a cluster of up to 16 CTAs
 * holds a whole tracer in its shared memory, so the rows cross HBM once (4 rows in, 1 out)
 * where the two-pass kernels read them twice. Applies on one rank with tree-ordered sums
 * when every tier-0 block has the fast shape, a CTA's slice (at most 8 blocks) fits five
 * row slots of shared memory and the cluster has at most 16 CTAs (86,400 cells at 675-leaf
 * blocks). Bit-identical to the two-pass kernels, which stay the default because they are
 * faster today. mode 1 = on where it applies, 0 = off (before
 * finish_setup); uses_cluster_caas reports the choice. */
int cedr_b200_set_cluster_caas(cedr_b200_cdr* cdr, int mode);
int cedr_b200_uses_cluster_caas(const cedr_b200_cdr* cdr, int* on);
int cedr_b200_uses_ring(const cedr_b200_cdr* cdr, int* on);
int cedr_b200_ring_info(const cedr_b200_cdr* cdr, int* info8_host);
/* Debug: per-unit, per-stage device timestamps of the last ring launch (only recorded when
 * the environment variable CEDR_B200_RING_TRACE is set); see tools/ring_trace.py. */
int cedr_b200_ring_trace(cedr_b200_cdr* cdr, unsigned long long* host, size_t cap, size_t* n);

/* Per-launch device times of run(): with profiling on, every kernel launch of
 * run() is bracketed by CUDA events on the CDR's stream (measurement aid for
 * bench.py's roofline line; off by default). After a synchronize,
 * cedr_b200_get_launch_times returns up to `cap` entries: ms_host[i] and a
 * kernel tag names_host[i] (see CEDR_B200_TAG_*). */
enum {
  CEDR_B200_TAG_RHOM = 0, CEDR_B200_TAG_UP = 1, CEDR_B200_TAG_TOP = 2,
  CEDR_B200_TAG_DOWN = 3, CEDR_B200_TAG_CAAS_ADJUST = 4, CEDR_B200_TAG_EXCHANGE = 5,
  CEDR_B200_TAG_FUSED = 6, CEDR_B200_TAG_MID = 7
};
int cedr_b200_set_profiling(cedr_b200_cdr* cdr, int on);
int cedr_b200_get_launch_times(cedr_b200_cdr* cdr, int cap, float* ms_host,
                               int* tags_host, int* tiers_host, int* n);
/* Debug builds only (-DCEDR_B200_PHASE_CLOCKS): clock sums per phase of the
 * specialised down-sweep, [0..8) leaf warp 0, [8..16) top warp; zeros otherwise. */
int cedr_b200_debug_phase_clocks(cedr_b200_cdr* cdr, unsigned long long* out16);
/* Tree plan facts: tiers, blocks in tier 0, max leaves per block, reference level
 * count (tree height + 1, cedr_tree.cpp:215-231). Any pointer may be NULL. */
int cedr_b200_plan_info(const cedr_b200_cdr* cdr, int* ntiers, int* nblocks0,
                        int* max_block_leaves, int* nlevels_ref);
/* Tuning knob, before finish_setup: maximum leaves per block (power users /
 * tests that want to force multi-tier plans on small trees). */
int cedr_b200_set_max_block_leaves(cedr_b200_cdr* cdr, int max_block_leaves);

/* Host-only probe of the tree plan (no device needed): flattens the tree, cuts it
 * into blocks and checks the plan the way the reference checks its comm plan
 * (tree::unittest -> test_comm_pattern, cedr_tree.cpp:279-348): the cell ids are
 * summed through the plan's block topology, tier by tier, and must give
 * ncells*(ncells-1)/2 at the root; every leaf and internal node must be used
 * exactly once. Outputs (any may be NULL): lci2gci_host[ncells], *ntiers,
 * nblocks_per_tier_host[<=64], *nshapes, *nlevels_ref, *idsum (the sum found). */
int cedr_b200_plan_probe(int ncells, int nnodes, int root, const int* kids,
                         const int64_t* cellidx, int max_block_leaves,
                         int64_t* lci2gci_host, int* ntiers,
                         int* nblocks_per_tier_host, int* nshapes, int* nlevels_ref,
                         int64_t* idsum);
/* tree::make_tree_over_1d_mesh as flat arrays (host): kids_host[2*(2*ncells-1)],
 * cellidx_host[2*ncells-1]; root is node 0. */
int cedr_b200_make_1d_tree(int ncells, int imbalanced, int* kids_host,
                           int64_t* cellidx_host);

/* Partial trees (tree::Node::level, cedr_tree_caller.hpp:20-22, consumed at
 * cedr_tree.cpp:71-76): a rank may hold the global tree with every subtree that contains
 * none of its cells cut down to a stub (a kid-less node of another rank,
 * cedr_tree.cpp:96-108). The block plan needs the whole tree on every rank, so the
 * parts are gathered once at setup and merged by this call (host only): part p is
 * part_nnodes[p] nodes at offset sum(part_nnodes[0..p)) of the concatenated flat arrays
 * (layout of cedr_b200_qlt_create; kids are part-local; internal nodes keep both kid
 * slots), part-local root part_root[p]; part p is rank p's. A position is internal if any
 * part expands it; a leaf's cell and rank come from its owner's part. The merged tree is
 * written in pre-order (root = node 0); pass null out arrays to query *out_nnodes. */
int cedr_b200_merge_partial_trees(int nparts, const int* part_nnodes, const int* part_root,
                                  const int* kids, const int64_t* cellidx,
                                  const int* node_rank, int cap_nodes, int* out_nnodes,
                                  int* out_kids, int64_t* out_cellidx, int* out_rank);
/* Setup-time helper: run the all-gather hook (device buffers) on `count` HOST doubles per
 * rank; recv_host holds nranks*count (rank-major). nranks == 1 copies. What Parallel's
 * MPI_Comm gives the reference for free at setup (cedr_mpi.hpp:17-27). */
int cedr_b200_allgather_host(cedr_b200_allgather_fn fn, void* ctx, int nranks,
                             const double* send_host, double* recv_host, size_t count);

/* Synthetic workload of SURVEY.md section 8(d), generated on the device:
 * splitmix64 stream seeded 0xCED20000 + config_id; fills rhom[ncells] and the
 * four [nt][lda] arrays. Bit-identical to compose_b200.workloads.headline(). */
int cedr_b200_fill_headline(int ncells, int nt, int64_t lda, int config_id,
                            double* rhom, double* qm_min, double* qm, double* qm_max,
                            double* qm_prev, void* cuda_stream);
/* The same workload restricted to cells [cell0, cell0 + nlclcells) (one rank's share):
 * rhom[nlclcells] and [nt][lda] arrays indexed by the local cell. */
int cedr_b200_fill_headline_range(int ncells, int cell0, int nlclcells, int nt, int64_t lda,
                                  int config_id, double* rhom, double* qm_min, double* qm,
                                  double* qm_max, double* qm_prev, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
