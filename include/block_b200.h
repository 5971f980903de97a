/* block_b200.h - C ABI of the B200-native DMRG sweep hot path (sigma = H.psi, Davidson, renormalisation).
 *
 * This is the drop-in boundary underneath the C++ entry points the CPU reference (sanshar/Block 1.1.1) exposes
 * to its sweep driver.  The reference has no FFI for this path: the seam is four C++ member
 * functions, so every entry point below cites the reference interface it replaces (file:line under the reference
 * root) and the C++ mirror in block_b200/host/ re-exposes them under the reference's own names on top of this ABI.
 *
 * Conventions
 *   - plain C, POD arguments only; every function returns 0 on success, non-zero on failure with the message
 *     available from b2d_last_error().  There is NO CPU fallback: a compute call on a context without a CUDA
 *     device fails with B2D_ERR_NO_DEVICE.
 *   - "side" is 0 for the left child of the big block, 1 for the right child (SpinBlock::get_leftBlock /
 *     get_rightBlock, spinblock.h:24).
 *   - sectors are (N, 2S, irrep) triples as in StateInfo::quanta (StateInfo.h:113); the point group is abelian
 *     (irrep product = XOR), which covers every BASELINE config.
 *   - host wavefunction buffers are FLAT in the reference's Wavefunction::FlattenInto order (wavefunction.C:167-186):
 *     allowed (lQ,rQ) blocks, lQ outer, each block row-major d_lQ x d_rQ.
 *   - host operator buffers are the allowed sector blocks of a SparseMatrix (BaseOperator.h:100-104) concatenated
 *     in (i outer, j inner) order, each block row-major d_i x d_j (newmat Matrix::Store(), newmat.h:455).
 *   - one host thread per context; one context per GPU (one process per GPU under torch.distributed / NCCL).
 */
#ifndef BLOCK_B200_H
#define BLOCK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2d_ctx b2d_ctx;

enum {
  B2D_OK = 0,
  B2D_ERR_ARG = 1,        /* bad argument / call order */
  B2D_ERR_NO_DEVICE = 2,  /* compute entry point called on a planning-only context (no CUDA device) */
  B2D_ERR_CUDA = 3,       /* a CUDA runtime call or kernel failed */
  B2D_ERR_NCCL = 4,
  B2D_ERR_NOCONV = 5      /* Davidson hit the iteration cap */
};

/* opTypes, BaseOperator.h:35-44 (only the ones the energy sweep carries, SURVEY.md 2.1 glossary) */
enum {
  B2D_HAM = 0, B2D_CRE = 1, B2D_CRE_CRE = 2, B2D_DES_DESCOMP = 3, B2D_CRE_DES = 4, B2D_CRE_DESCOMP = 5,
  B2D_CRE_CRE_DESCOMP = 6, B2D_OVERLAP = 13
};

/* ---- context -------------------------------------------------------------------------------------------- */

/* device >= 0: bind to that CUDA device.  device == -1: planning-only context (integer/host work only; used by
 * the CPU tests; every compute call fails with B2D_ERR_NO_DEVICE). */
int b2d_create(int device, b2d_ctx** out);
void b2d_destroy(b2d_ctx* ctx);
/* Forget the block description (children, operators, plan, wavefunction slots, density / rotation state) and keep what is
 * expensive to create (streams, pinned memory, library handles, operator arena slabs, scratch buffers): a sweep creates one
 * context and resets it between block iterations (the reference builds a new `big` SpinBlock per block iteration, sweep.C:182). */
int b2d_reset(b2d_ctx* ctx);
const char* b2d_last_error(const b2d_ctx* ctx);   /* ctx may be NULL: last error of a failed b2d_create */
int b2d_abi_version(void);                        /* 4 */

/* Tuning knobs (all optional): "workspace_mb" (T workspace for the two-step contraction), "max_davidson_iter",
 * "tile_class" (debug: -1 auto, 0/1/2 = square 128/64/32 DMMA tiles everywhere, 3 = auto with the tiny-sector warp kernel,
 * which auto already uses), "sync_debug", "phase_timing", "opbuild_batch" (b2d_build_enlarged_op defers its scatter tasks and
 * b2d_stash_product / b2d_product_op_download run them for the whole block, one launch per round instead of one per product; same
 * summation order; default on), "factorised" (operators of an enlarged block whose right child is a one-site dot are NOT materialised:
 * b2d_build_enlarged_op / b2d_product_op_accumulate record, per sector block, the list of scaled sub-blocks of the renormalised child's
 * operators - operatorfunctions.C:188-250 rowstride / colstride structure - and sigma, diag(H), the noise products and the operator
 * rotation contract those factors directly: ~16x less operator memory, no flops on the structural zeros of the Kronecker blocks),
 * "eig_jacobi_max" (largest sector for the single-CTA Jacobi kernel, default 64; larger sectors use the block Jacobi kernel),
 * "eig_cusolver" (diagnostic: cusolverDnDsyevd for the large sectors, never the default), "slice_iters" (pipeline iterations per
 * split-K slice of a sigma block, default 256, at most 8 slices), "slice_iters_narrow" (> 0: the narrow tiles of a sigma block - the
 * remainder bands of ragged sectors - get their own, finer slicing, at most 32 slices; default 0 = off: measured +0.7 % on the
 * benchmark for 2 GB of partial buffers), "presum_identity", "balance_terms" (cost-weighted term ownership for several ranks),
 * "cache_device_mb" (device budget of the block cache; <= 0: automatic), "partition_renormalisation" (several ranks that ALL hold the
 * whole block - the multi-process drop-in - divide the sectors of b2d_diagonalise_dm and the operators of b2d_transform_operators by
 * cost and all-reduce the zero-filled results: every rank ends with identical bits; every rank must set it). */
int b2d_set_option(b2d_ctx* ctx, const char* key, double value);

/* ---- block description: replaces the host-side SpinBlock / StateInfo / Op_component objects ---------------- */

/* StateInfo of one child (StateInfo.h:113-147: quanta, quantaStates) + SpinBlock::sites / is_loopblock()
 * (spinblock.h:104,138). q = nq x 3 ints (N, 2S, irrep). */
int b2d_set_block(b2d_ctx* ctx, int side, int nq, const int32_t* q, const int32_t* dims,
                  int is_loop, int nsites, const int32_t* sites);

/* One SparseMatrix (BaseOperator.h:75-243) of an operator array of that child (Op_component<Op>,
 * op_components.h:214): optype, orbital indices (norb = 0,1,2), spin-component index inside the vector returned by
 * get_element (op_components.C:172-184), deltaQuantum[0], fermion flag, allowedQuantaMatrix (nq x nq bytes) and the
 * packed blocks.  data may be NULL: the blocks are allocated zero-filled on the device (see b2d_fill_op_random).
 * Returns the operator id (>= 0) in *op_id. */
int b2d_add_op(b2d_ctx* ctx, int side, int optype, int norb, const int32_t* orbs, int comp, const int32_t* dq,
               int fermion, const uint8_t* allowed, const double* data, int* op_id);

/* The same with one pointer per allowed block (i outer, j inner; each block row-major d_i x d_j, contiguous - newmat Matrix::Store()):
 * the library copies the blocks straight from the caller's matrices into its pinned staging buffer, so a binding needs no packed
 * intermediate copy.  Uploads are batched: the device copy happens when 64 MB are pending or at b2d_plan. */
int b2d_add_op_blocks(b2d_ctx* ctx, int side, int optype, int norb, const int32_t* orbs, int comp, const int32_t* dq,
                      int fermion, const uint8_t* allowed, const double* const* blocks, int* op_id);

/* Synthetic-benchmark helper: fill an operator's device blocks with a counter-based uniform(-a,a) stream; if
 * symmetric != 0 the operator is made self-adjoint in the reduced-matrix-element sense (needs dq = (0,0,0)). */
int b2d_fill_op_random(b2d_ctx* ctx, int side, int op_id, uint64_t seed, double amplitude, int symmetric);

/* Allocate (zero-filled) every operator of a side that was added with data == NULL: needed for the children of a product block
 * (b2d_set_product_stateinfo), which never go through b2d_plan. */
int b2d_alloc_ops(b2d_ctx* ctx, int side);

/* Copy an operator's blocks back in the host layout of b2d_add_op. */
int b2d_download_op(b2d_ctx* ctx, int side, int op_id, double* data);
int64_t b2d_op_size(const b2d_ctx* ctx, int side, int op_id);   /* packed doubles */

/* Wavefunction target quantum (Wavefunction::initialise, wavefunction.C:18-56), core energy (coreEnergy[],
 * spinblock.C:735), Hamiltonian type (HUBBARD skips the two-index terms, spinblock.C:771), number of spatial
 * orbitals (the `length` of trimap_2d, para_array.h:360) and this process's share of the operator terms
 * (rank, nranks: the partition of distribute.C / para_array.h:33-42 across GPUs).
 * Builds the psi layout, the term list of SpinBlock::multiplyH (spinblock.C:722-789) with every scalar factor
 * (9j, parities, transpose scalings; operatorfunctions.C:520-529) and the grouped-contraction schedule. */
int b2d_plan(b2d_ctx* ctx, const int32_t* psi_dq, double core_energy, int hubbard, int norbs, int rank, int nranks);

int64_t b2d_psi_size(const b2d_ctx* ctx);          /* W: length of a flat host wavefunction */
int64_t b2d_psi_padded_size(const b2d_ctx* ctx);   /* device-internal padded length */
int b2d_psi_num_blocks(const b2d_ctx* ctx);
/* per allowed block: lQ, rQ, flat offset (== StateInfo::unBlockedIndex of the big block, StateInfo.C:213-227) */
int b2d_psi_blocks(const b2d_ctx* ctx, int32_t* lq, int32_t* rq, int64_t* offset);

/* Term list introspection (integer work, bit-exact contract).  One row per TensorMultiply call of
 * multiplyH that THIS rank executes: left op id, right op id, transpose flags (bit0 = left is a Transposeview,
 * bit1 = right), scale, owner rank.  all_ranks != 0 lists every rank's terms (owner column tells whose). */
int b2d_num_terms(const b2d_ctx* ctx, int all_ranks);
int b2d_terms(const b2d_ctx* ctx, int all_ranks, int32_t* left_op, int32_t* right_op, int32_t* flags,
              double* scale, int32_t* owner);
/* ALGORITHMIC flops of one multiplyH = the dgemm flops the reference issues (operatorfunctions.C:515,530);
 * all_ranks = 0: this rank's share. */
double b2d_sigma_flops(const b2d_ctx* ctx, int all_ranks);
/* schedule statistics: out[0]=#chunks, [1]=#step-1 contractions, [2]=#step-2 segments, [3]=#sigma tiles,
 * [4]=workspace doubles, [5]=operator arena doubles, [6]=kernel launches per sigma, [7]=flops executed,
 * [8]=useful flops inside tiles, [9]=flops the tiles issue including ragged-edge padding, [10]=doubles of pre-summed factor blocks
 * ("combos") of factorised operators, [11]=factors that point straight at a child operator's block, [12]=factors that point at a combo */
int b2d_plan_stats(const b2d_ctx* ctx, double* out, int n);

/* ---- device-resident wavefunction slots ------------------------------------------------------------------- */

int b2d_vec_reserve(b2d_ctx* ctx, int nslots);                       /* slots 0..nslots-1, zero-filled */
int b2d_vec_upload(b2d_ctx* ctx, int slot, const double* flat);      /* Wavefunction::CollectFrom, wavefunction.C:188 */
int b2d_vec_download(b2d_ctx* ctx, int slot, double* flat);          /* Wavefunction::FlattenInto, wavefunction.C:167 */
int b2d_vec_dot(b2d_ctx* ctx, int a, int b, double* out);            /* DotProduct, BaseOperator.C:240 */
int b2d_vec_axpy(b2d_ctx* ctx, double alpha, int x, int y);          /* ScaleAdd, BaseOperator.C:255: y += alpha x */
int b2d_vec_scale(b2d_ctx* ctx, double alpha, int x);                /* Scale, BaseOperator.C:272 */
int b2d_vec_copy(b2d_ctx* ctx, int src, int dst);
int b2d_vec_clear(b2d_ctx* ctx, int slot);

/* ---- sigma ------------------------------------------------------------------------------------------------ */

/* SpinBlock::multiplyH(Wavefunction& c, Wavefunction* v, int) spinblock.h:235, spinblock.C:722-789.
 * dst (+)= H src over this rank's terms, followed (nranks > 1, communicator attached) by the all-reduce that
 * replaces distributedaccumulate (distribute.h:42-76).  accumulate = 1 keeps the reference's "v += H c"
 * contract; 0 overwrites. */
int b2d_sigma(b2d_ctx* ctx, int src_slot, int dst_slot, int accumulate);

/* The same call with HOST buffers (what a reference-side binding does per Davidson_functor call,
 * davidson.C:19-22): H2D of c, sigma, D2H of v.  accumulate = 1: v += H c exactly as the reference (v is uploaded
 * too); accumulate = 0: v = H c (what block_davidson needs: it clears v first, linear.C:239-240). */
int b2d_multiplyH_host(b2d_ctx* ctx, const double* c_flat, double* v_flat, int accumulate);

/* operatorfunctions::TensorMultiply(ablock, a, b, cblock, c, v, opQ, scale), operatorfunctions.C:485-537, for ONE
 * operator pair (left op id, right op id, transpose flags as in b2d_terms): dst += scale (A_L x A_R) src.
 * opq_spin is the 2S of opQ (0 for every Hamiltonian term).  right_op < 0 or left_op < 0 selects the
 * one-operator form (operatorfunctions.C:331-404) with the identity on the missing side. */
int b2d_tensor_multiply(b2d_ctx* ctx, int left_op, int right_op, int flags, int opq_spin, double scale,
                        int src_slot, int dst_slot);

/* One-operator form, operatorfunctions::TensorMultiply(ablock, a, cblock, c, v, dQ, scale) operatorfunctions.C:331-404:
 * v += scale (a x 1) c (side 0) or scale (1 x a) c (side 1; fermion sign of the left sector included, :393), with HOST
 * buffers.  c is a wavefunction of the planned target quantum; v is a wavefunction with target quantum dst_dq (the
 * operator shifts the sector: Wavefunction(q, &big, onedot), density.C:215), flat in ITS FlattenInto order,
 * b2d_wavefunction_size(ctx, dst_dq) doubles.  transposed != 0 applies the Transposeview of the operator. */
int64_t b2d_wavefunction_size(b2d_ctx* ctx, const int32_t* dq);
int b2d_tensor_multiply_one_host(b2d_ctx* ctx, int side, int op_id, int transposed, const int32_t* dst_dq, double scale, const double* c_flat,
                                 double* v_flat);

/* SpinBlock::diagonalH(DiagonalMatrix&) spinblock.h:240, spinblock.C:855-899: diag(H) in flat psi order. */
int b2d_diagonal(b2d_ctx* ctx, int dst_slot);

/* ---- Davidson --------------------------------------------------------------------------------------------- */

/* Linear::block_davidson(vector<Wavefunction>& b, DiagonalMatrix& h_diag, double normtol, const bool& warmUp,
 * Davidson_functor& h_multiply, bool& useprecond, int currentRoot, vector<Wavefunction>& lowerStates)
 * linear.h:28, linear.C:179-385, state-averaged form (currentRoot = -1, no lower states).
 * Guesses are in slots guess_slot0 .. guess_slot0+nroots-1 and are overwritten by the solutions (as `b` is).
 * diag_slot holds diag(H).  evals[nroots] receives the eigenvalues (h_diag.element(i), linear.C:344),
 * *n_multiply the number of H applications, *residual the last ||r||^2.
 * The Krylov vectors, sigma vectors, subspace matrix, its eigen-decomposition and the Olsen preconditioner all
 * stay on the device; the host only reads one convergence scalar per iteration.  The reference has no iteration cap
 * (linear.C:214 `maxiter` is unused); the option "max_davidson_iter" (default 2000) stops a runaway solve: the current
 * Ritz pairs are returned together with B2D_ERR_NOCONV.
 * Limits (B2D_ERR_ARG otherwise): 1 <= nroots <= deflation_min < deflation_max <= 30 - the subspace matrix, its eigenvectors and the
 * fused multi-vector kernels are sized for 32 vectors.  The reference accepts any deflation_min_size / deflation_max_size from its input
 * file (input.C); a binding keeps such runs working by leaving block_davidson to the reference and routing only its H applications here
 * (b2d_multiplyH_host) - tests/dropin/block_gpu_hooks.cpp does exactly that. */
int b2d_davidson(b2d_ctx* ctx, int nroots, int guess_slot0, int diag_slot, double normtol, int deflation_min,
                 int deflation_max, double* evals, int* n_multiply, double* residual);
/* The same solve with `lowerStates` (state-specific form, currentRoot >= 0; linear.C:201-208, 311-317, 369-375): the first guess,
 * every residual (before its norm is taken) and every new Krylov vector are projected against the n_lower wavefunctions in slots
 * lower_slot0.. as r <- r - <r|l>/<l|l> l, in the order given.  The caller has orthogonalised the lower states among
 * themselves (solver.C:79-86).  n_lower = 0 is b2d_davidson. */
int b2d_davidson_lower(b2d_ctx* ctx, int nroots, int guess_slot0, int diag_slot, double normtol, int deflation_min,
                       int deflation_max, int n_lower, int lower_slot0, double* evals, int* n_multiply, double* residual);

/* ---- renormalisation ---------------------------------------------------------------------------------------- */

/* DensityMatrix::makedensitymatrix (density.C:27-90, noise = 0): rho[q] = sum_i w_i sum_r psi_i[q,r] psi_i[q,r]^T
 * for the wavefunctions in slots slot0..slot0+nroots-1. */
int b2d_make_density(b2d_ctx* ctx, int nroots, int slot0, const double* weights);
/* DensityMatrix::add_onedot_noise (density.C:332-399, functor onedot_noise_f :181-258) for every root, as makedensitymatrix
 * does when the schedule's noise > 0 (density.C:40-60): rho += (noise/nroots) / tr(rho_n) * rho_n with
 * rho_n = sum_O (O psi)(O psi)^T / |O psi|^2 over the left block's CRE, CRE_CRE, CRE_DES (or DES_DESCOMP, CRE_DESCOMP) operators
 * and their transposes, each into its +-dQ shifted sector.  Under a term partition every rank does its operators and rho_n
 * is all-reduced.  Call after b2d_make_density.  (additional_noise / add_twodot_noise: b2d_add_wavefunction_density.) */
int b2d_add_onedot_noise(b2d_ctx* ctx, int nroots, int slot0, double noise);
/* rho += weight * w w^T for a HOST wavefunction w of any target quantum dq, flat in its own FlattenInto order
 * (b2d_wavefunction_size(ctx, dq) doubles): MultiplyProduct(w, Transpose(w), dm, weight), operatorfunctions.C:630-650.  This is the
 * device half of DensityMatrix::add_twodot_noise (density.C:92-165): the caller draws and normalises the random wavefunctions
 * (Wavefunction::Randomise, glibc rand()) so that the random stream is the reference's own.  Call after b2d_make_density. */
int b2d_add_wavefunction_density(b2d_ctx* ctx, const int32_t* dq, const double* flat, double weight);
int64_t b2d_density_size(const b2d_ctx* ctx);       /* sum_q d_q^2 */
int b2d_density_download(b2d_ctx* ctx, double* rho);   /* blocks q = 0..nq-1, row-major d_q x d_q */
int b2d_density_upload(b2d_ctx* ctx, const double* rho);

/* diagonalise_dm (rotationmat.C:258-279): per-sector symmetric eigen-decomposition on the device, eigenvalues
 * ascending, those < 1e-14 set to 0.  evals receives sum_q d_q doubles. */
int b2d_diagonalise_dm(b2d_ctx* ctx, double* evals);

/* sort_weights + assign_matrix_by_dm (rotationmat.C:313-346, :149-256; keptqstates = 0): global descending order
 * with the reference's tie rule, keep the first min(total, keep_states) with weight > 1e-13.
 * kept_counts[nq] receives the retained states per left sector, *discarded the discarded weight.  The rotation
 * matrices (kept eigenvectors as columns, selection order) stay on the device. */
int b2d_select_states(b2d_ctx* ctx, int keep_states, int32_t* kept_counts, double* discarded);

int64_t b2d_rotation_size(const b2d_ctx* ctx);                 /* sum_q d_q * kept_q */
int b2d_rotation_download(b2d_ctx* ctx, double* rot);          /* per sector, row-major d_q x kept_q */
int b2d_rotation_upload(b2d_ctx* ctx, const int32_t* kept_counts, const double* rot);

/* SpinBlock::transform_operators(vector<Matrix>&) spinblock.h:253, save_load_block.C:267-319 ->
 * SparseMatrix::renormalise_transform BaseOperator.C:341-363 -> MatrixRotate MatrixBLAS.C:553-572:
 * every operator of the LEFT child becomes O'[a,b] = U_Q(a)^T O[Q(a),Q(b)] U_Q(b) on the retained sectors.
 * The rotated operators replace nothing: they are written into a fresh arena that b2d_rotated_* reads. */
int b2d_transform_operators(b2d_ctx* ctx);
int b2d_rotated_num_sectors(const b2d_ctx* ctx);
int b2d_rotated_sectors(const b2d_ctx* ctx, int32_t* old_index, int32_t* dims);
int64_t b2d_rotated_op_size(const b2d_ctx* ctx, int op_id);
int b2d_rotated_op_download(b2d_ctx* ctx, int op_id, uint8_t* allowed, double* data);   /* data may be NULL: mask only */
/* All rotated operators in one device pass and one copy: the packed blocks of operator 0, 1, ... back to back
 * (b2d_rotated_op_size(id) doubles each, b2d_rotated_total_size in total). */
int64_t b2d_rotated_total_size(const b2d_ctx* ctx);
int b2d_rotated_download_all(b2d_ctx* ctx, double* data);

/* SpinBlock::RenormaliseFrom (spinblock.h:247-251, renormalise.C:39-133), two-dot: diagonalH, Davidson from the guesses in
 * slot0.., density matrix (+ one-dot noise if noise > 0), eigen-decomposition, state selection.  Leaves the solutions in
 * the guess slots and the rotation matrices on the device (follow with b2d_transform_operators). */
int b2d_renormalise_from(b2d_ctx* ctx, int nroots, int guess_slot0, const double* weights, double normtol,
                         int keep_states, int deflation_min, int deflation_max, double noise, double* energies,
                         int32_t* kept_counts, double* discarded, int* n_multiply);

/* ---- construction of enlarged-block operators on the device (first device step) --------------------
 * operatorfunctions::TensorProduct(ablock, a, b, cblock, cstateinfo, c, scale) operatorfunctions.C:146-254 and
 * operatorfunctions::TensorTrace operatorfunctions.C:19-117 (-> MatrixTensorProduct MatrixBLAS.C:125-200): the scatter every
 * Op::build of an enlarged block is made of (Operators.C:453-2395).  Here side 0 / side 1 of the context are the two CHILDREN of the
 * enlarged block (renormalised block and dot), described with b2d_set_block / b2d_add_op as usual; no b2d_plan is needed.
 * b2d_product_op_* are the primitives; b2d_build_enlarged_op plans and builds a whole operator. */

/* Product StateInfo of the enlarged block after CollectQuanta (StateInfo.h:113-147): collected quanta q (nq x 3) and sizes;
 * per UNCOLLECTED sector u its left / right child sectors (leftUnMapQuanta, rightUnMapQuanta) and size
 * (unCollectedStateInfo->quantaStates); per collected sector c its pieces old_to_new[old_to_new_begin[c] .. old_to_new_begin[c+1])
 * (oldToNewState), in the order in which they are concatenated. */
int b2d_set_product_stateinfo(b2d_ctx* ctx, int nq, const int32_t* q, const int32_t* dims, int nunc, const int32_t* lmap,
                              const int32_t* rmap, const int32_t* unc_dims, const int32_t* old_to_new_begin,
                              const int32_t* old_to_new);
/* SparseMatrix::allocate(stateinfo) BaseOperator.C:123-145 for an operator with deltaQuantum dq on the enlarged block: zero-filled
 * device blocks where q_i is in dq (+) q_j. */
int b2d_product_op_create(b2d_ctx* ctx, const int32_t* dq, int fermion, int* prod_id);
/* c += scale * (a x b), a = operator left_op of the left child (or its Transposeview), b = operator right_op of the right child;
 * left_op < 0 or right_op < 0: the identity on that child (TensorTrace).  9j coefficient, Transposeview scalings and the fermion
 * sign of operatorfunctions.C:205-218 are applied per sub-block. */
int b2d_product_op_accumulate(b2d_ctx* ctx, int prod_id, int left_op, int left_transposed, int right_op, int right_transposed,
                              double scale);
/* One- and two-electron integrals as the reference's accessors return them for the (reordered) spatial orbitals:
 * v1[i*n + j] = v_1(2i, 2j), v2[((i*n + j)*n + k)*n + l] = v_2(2i, 2j, 2k, 2l) (IntegralMatrix.C:32-46, 312-333), the irrep of every
 * spatial orbital and the screening thresholds (input.C:141-142).  Kept across b2d_reset. */
int b2d_set_integrals(b2d_ctx* ctx, int norbs, const double* v1, const double* v2, const int32_t* orbital_irreps, double one_tol,
                      double two_tol);
/* Op::build of ONE operator of the enlarged block (Operators.C:453-2625: Cre, CreCre, CreDes, CreDesComp, DesDesComp, CreCreDesComp,
 * Ham, Overlap; energy sweep, spin-adapted, abelian): the host planner (block_b200/csrc/opbuild.hpp: TensorOp coupling,
 * calcCompfactor with the integrals, commute parities, 6j recoupling) decides which child products enter it, the device
 * performs them.  b2d_enlarged_op_products only lists them (integer / scalar work, no device needed; flags bit0 = left operand is a
 * Transposeview, bit1 = right; an id of -1 = identity on that child): returns the count, or -(error code). */
int b2d_enlarged_op_products(b2d_ctx* ctx, int optype, int norb, const int32_t* orbs, const int32_t* dq, int hubbard, int max_products,
                             int32_t* left_op, int32_t* right_op, int32_t* flags, double* scale);
int b2d_build_enlarged_op(b2d_ctx* ctx, int optype, int norb, const int32_t* orbs, int comp, const int32_t* dq, int fermion,
                          int hubbard, int* prod_id);
/* Both children of the big block of a two-dot step are themselves enlarged blocks (SpinBlock::BuildSumBlock).  A child that was
 * built on the device from ITS children (b2d_build_enlarged_op for each of its operators, in the order of the reference's operator
 * arrays) is parked with b2d_stash_product (is_loop / sites as in b2d_set_block); a child that was uploaded as it is (b2d_set_block +
 * b2d_add_op on side `from_side`) with b2d_stash_side; b2d_assemble_big makes the block parked in slot 0 / 1 the left / right child,
 * after which b2d_plan runs as usual.  Operators built on the device never cross the PCIe bus. */
int b2d_stash_product(b2d_ctx* ctx, int slot, int is_loop, int nsites, const int32_t* sites);
int b2d_stash_side(b2d_ctx* ctx, int slot, int from_side);
int b2d_assemble_big(b2d_ctx* ctx);
/* Write a FACTORISED operator (option "factorised") of child `side` of the assembled big block out as dense sector blocks, in place; it is
 * an ordinary materialised operator afterwards.  The benchmark uses it to run the same operator pair through both forms at full size. */
int b2d_materialise_op(b2d_ctx* ctx, int side, int op_id);
/* Measurement: out[0..3] = {products, scatter tasks, ALGORITHMIC bytes of those tasks (8 x (|A| + |B| + 2 |destination piece|): operands read
 * once, destination read-modified-written), launches of the last batched flush} since b2d_set_product_stateinfo; with option opbuild_batch
 * b2d_last_timing gives the CUDA-event time of the batched flush. */
int b2d_product_stats(const b2d_ctx* ctx, double* out, int n);
int64_t b2d_product_op_size(const b2d_ctx* ctx, int prod_id);
int b2d_product_op_download(b2d_ctx* ctx, int prod_id, uint8_t* allowed, double* data);   /* host layout of b2d_add_op */

/* ---- guess wavefunction of the next block iteration ----------------------------------------- */

/* One StateInfo of the reference as the transform reads it (StateInfo.h:60-130): quanta / quantaStates; newQuantaMap for a
 * renormalised block (index of each sector in the un-truncated StateInfo it was cut from; NULL otherwise); for a collected
 * product StateInfo the unCollectedStateInfo tables (quanta, quantaStates, leftUnMapQuanta, rightUnMapQuanta) and oldToNewState
 * in CSR form (old_to_new_begin has nq + 1 entries); nunc = 0 and NULLs otherwise. */
typedef struct b2d_stateinfo {
  int32_t nq;
  const int32_t* q;
  const int32_t* dims;
  const int32_t* new_quanta_map;
  int32_t nunc;
  const int32_t* unc_q;
  const int32_t* unc_dims;
  const int32_t* unc_left;
  const int32_t* unc_right;
  const int32_t* old_to_new_begin;
  const int32_t* old_to_new;
} b2d_stateinfo;

/* Everything GuessWave::transform_previous_wavefunction (guess_wavefunction.C:524-636, two-dot branch) reads:
 *   dq        target quantum of the wavefunction
 *   sys       big.leftStateInfo->leftStateInfo   the renormalised system block S' (newQuantaMap -> sectors of `oldleft`)
 *   dot       big.leftStateInfo->rightStateInfo  the new system dot (one state per sector)
 *   left      big.leftStateInfo                  S' (x) dot, collected, with its un-collected tables
 *   right     big.rightStateInfo                 the new environment side (sector dims only)
 *   oldleft   oldStateInfo.leftStateInfo         row space of the previous wavefunction (dims only)
 *   oldright  oldStateInfo.rightStateInfo        its column space E_old (x) dot, collected, with its un-collected tables
 *   env       oldStateInfo.rightStateInfo->leftStateInfo   E_old (newQuantaMap -> sectors of `right`)
 *   old_allowed  oldleft.nq x oldright.nq: allowed blocks of the previous wavefunction (wave-*.tmp)
 *   lrot_cols / rrot_cols  kept states per sector of the two rotation matrices (Rotation-*.tmp; 0 = sector dropped): the left one
 *                          is indexed by `oldleft` sectors, the right one by `right` sectors
 *
 * One-dot steps (GuessWave::onedot_transform_wavefunction, guess_wavefunction.C:832-936): the previous wavefunction is [S.d][E_old]; its
 * columns are first expanded with the right rotation matrix, its rows rotated with the left one.
 *   mode 1  dot on the system side (transpose_guess_wave): sys / dot / left / right as above; oldcol = oldStateInfo.rightStateInfo
 *           (newQuantaMap -> sectors of `oldright`); oldright = right (x) dot, collected (the reference's newenvstateinfo, :853-856) with
 *           its un-collected tables; env unused.  The right rotation matrix (sites of the right block + the dot) is indexed by `oldright`
 *           sectors; the rotated wavefunction [S'][E'.d] is shuffled to [S'.d][E'].
 *   mode 2  dot on the environment side: left = big.leftStateInfo (newQuantaMap -> sectors of `oldleft`), right, oldleft, oldcol
 *           (newQuantaMap -> sectors of `right`); no shuffle.  sys, dot, oldright, env unused.
 *   mode 3  TRANSPOSE guess of the first block iteration of a sweep (GuessWave::transpose_previous_wavefunction, :55-84, two-dot to
 *           two-dot): left, right, oldleft (= the sectors of `right`), oldcol (= the sectors of `left`); trial(i, j) = parity . old(j, i)^T;
 *           no rotation matrices (pass lrot_cols / rrot_cols = NULL).
 *   mode 4  TRANSPOSE guess of the first block iteration of a ONE-DOT sweep (GuessWave::onedot_transpose_wavefunction, :140-198):
 *           [S.d][E] -> [E.d][S]; left (E (x) d, with its un-collected tables), sys (= E), dot, right (= S), oldleft (S (x) d collected, with
 *           its un-collected tables: left -> sectors of `right`, right -> dot sectors), oldcol (= the sectors of `sys`); no rotation matrices.
 * old_allowed is oldleft.nq x oldcol.nq in modes 1 to 4. */
typedef struct b2d_guess_desc {
  int32_t dq[3];
  int32_t mode;     /* 0 two-dot, 1 / 2 one-dot, 3 / 4 transpose (see above) */
  b2d_stateinfo sys, dot, left, right, oldleft, oldright, env, oldcol;
  const uint8_t* old_allowed;
  const int32_t* lrot_cols;
  const int32_t* rrot_cols;
} b2d_guess_desc;

/* Plan the transform (host integer / scalar work: sector bookkeeping, getCommuteParity, 6j recoupling coefficients; valid on a
 * planning-only context).  out[0..7] = {doubles of the previous wavefunction, of the left rotation, of the right rotation, of
 * the trial vector (flat), GEMM flops, algorithmic bytes of the shuffle, number of shuffle tasks, number of shuffle rounds}. */
int b2d_guess_plan(b2d_ctx* ctx, const b2d_guess_desc* desc, double* out, int n);
/* The plan's descriptors for inspection (CPU tests execute them with numpy), in execution order: what = 0 / 1 segments (GSeg, 40 bytes
 * each) / groups (GGroup, 40) of the first contraction batch, 10 / 11 of the second (one-dot only), 2 shuffle tasks (KronTask, 80; rounds
 * concatenated; pad bit 0: the destination offset is into the trial vector, bit 1: the source offset is into the input image), 3 tasks per round (int32), 4 / 5 segments / groups of the batch
 * after the shuffle (two-dot only), 6 input blocks (BlockDesc, 32: previous wavefunction, left rotation, right rotation), 7 counts of
 * those three tables (int32 x 3) followed by {image, T1, work} sizes in doubles (int64 x 3 at byte 16), 8 trial blocks (BlockDesc).
 * Returns the number of bytes (copied when cap is large enough), or -1. */
int64_t b2d_guess_plan_export(const b2d_ctx* ctx, int what, void* out, int64_t cap);
/* Execute it on the device: the three host arrays (allowed blocks row-major, (i outer, j inner) order; rotation matrices as the
 * reference stores them, d_q x m_q row-major, dropped sectors skipped) are uploaded once, stage 1 / 3 run in the grouped FP64
 * contraction kernel, stage 2 in the HBM-bound scatter kernel.  dst_slot >= 0: the trial vector is left in that wavefunction slot
 * (needs b2d_plan with the same children and target quantum: the layouts must be identical) ready for b2d_davidson, nothing
 * returns to the host; trial != NULL: it is (also) downloaded in FlattenInto order. */
int b2d_guess_transform(b2d_ctx* ctx, const double* old_wave, const double* left_rot, const double* right_rot, int dst_slot, double* trial);

/* ---- device-side shadow of the scratch files --------------------------------------------------
 * The reference writes every renormalised block to a scratch file (SpinBlock::store, save_load_block.C:23-64, called at sweep.C:466) and
 * reads the environment block back (SpinBlock::restore :66-108, initblocks.C:279-288) each block iteration.  Here the renormalised
 * operators never leave the GPU: b2d_cache_put_rotated keeps the block b2d_transform_operators just produced under a token (the buffer
 * changes owner, nothing is copied); b2d_cache_use makes a cached block child `side` of the next product (b2d_set_product_stateinfo) or
 * big block (b2d_plan) in place of b2d_set_block + b2d_add_op.  The binding stores the token where the reference stores the matrices (see
 * INTEGRATION.md), so it travels through the reference's own store / restore / copies.  Device memory beyond "cache_device_mb" (option;
 * default: automatic - entries stay on the device while 35 % of the GPU's memory remains free) spills to pinned host memory.  b2d_cache_download_op returns an operator in the host layout of b2d_add_op when a host
 * code path needs the matrices after all; entries live until b2d_cache_drop / b2d_destroy (b2d_reset keeps them). */
int b2d_cache_put_rotated(b2d_ctx* ctx, uint64_t* token);
int b2d_cache_use(b2d_ctx* ctx, uint64_t token, int side, int is_loop);
int b2d_cache_block_info(const b2d_ctx* ctx, uint64_t token, int32_t* nq, int32_t* nops, int32_t* nsites);
int b2d_cache_block_sectors(const b2d_ctx* ctx, uint64_t token, int32_t* q, int32_t* dims, int32_t* sites);
int b2d_cache_op_info(const b2d_ctx* ctx, uint64_t token, int op_id, int32_t* optype, int32_t* norb, int32_t* orbs, int32_t* comp, int64_t* packed_size);
int b2d_cache_download_op(b2d_ctx* ctx, uint64_t token, int op_id, uint8_t* allowed, double* data);
int b2d_cache_drop(b2d_ctx* ctx, uint64_t token);
/* out[0..5] = {entries, doubles held on the device, doubles spilled to pinned host memory, puts, uses, evictions} */
int b2d_cache_stats(const b2d_ctx* ctx, double* out, int n);
/* Memory pressure (the role of the scratch DISK of save_load_block.C:23-108 when the blocks of a sweep exceed the memory: P5 at M = 4000
 * stores 20 - 60 GB per block).  Every device allocation of the library that fails first releases the empty arena slabs and spare
 * buffers and then moves cached blocks that are not children of the current block iteration (oldest first) to pinned host memory, and
 * tries again.  b2d_cache_spill asks for the same explicitly until `bytes` of device memory are free (a host that needs GPU memory for
 * itself; the tests).  Returns B2D_OK also when less could be freed: b2d_cache_stats tells what moved. */
int b2d_cache_spill(b2d_ctx* ctx, double bytes);

/* ---- multi-GPU: partition of operator terms, NCCL all-reduce of the partial sigma --------------------------- */

int b2d_nccl_unique_id(uint8_t* id128);                                   /* rank 0; broadcast out of band */
int b2d_comm_init(b2d_ctx* ctx, const uint8_t* id128, int rank, int nranks);
int b2d_allreduce_slot(b2d_ctx* ctx, int slot);

/* ---- measurement ------------------------------------------------------------------------------------------- */

/* CUDA-event time (ms) of the last b2d_sigma / b2d_davidson / b2d_make_density / b2d_transform_operators call on
 * the context's stream: out[0] = total, out[1] = step-1 kernels, out[2] = step-2 kernels, out[3] = collective. */
int b2d_last_timing(b2d_ctx* ctx, double* out, int n);
/* One multiplyH (this rank's terms, no collective) with CUDA events around every launch of the grouped contraction
 * kernel.  out[(step * 10 + tile_class) * 4 + {0,1,2,3}] = {summed kernel ms, useful flops (2mnk) executed, flops the
 * tiles issue including ragged-edge padding, number of launches}; step 0 = T = A_L psi (operatorfunctions.C:512-516),
 * step 1 = sigma += F T A_R^T (:517-531); tile class = 3 * r + c for a (128 >> r) x (128 >> c) DMMA tile, class 9 = the
 * warp-per-block DFMA + shuffle-reduction kernel of the tiny (<= 8 x 8) sectors.  80 doubles. */
int b2d_sigma_profile(b2d_ctx* ctx, int src_slot, int dst_slot, double* out);
int64_t b2d_kernel_launches(const b2d_ctx* ctx);     /* kernels launched by this context so far */
int b2d_sync(b2d_ctx* ctx);
void* b2d_stream(b2d_ctx* ctx);                      /* cudaStream_t, for event timing by the caller */
/* FP64 yardsticks measured with this library's own kernels on the context's device: DMMA register loop
 * (TFLOP/s), DFMA register loop (TFLOP/s). */
int b2d_measure_fp64_peak(b2d_ctx* ctx, double* dmma_tflops, double* dfma_tflops);

/* HBM-bound level-1 kernels of the Davidson iteration (DotProduct / ScaleAdd / Normalise of BaseOperator.C:240-288 and the Ritz
 * rotation of linear.C:279-294 in their fused device forms) timed alone with CUDA events on wavefunction slots slot0..slot0+17
 * (overwritten).  out[2k] = ms per launch, out[2k+1] = algorithmic bytes per launch; k = 0 multi_dot(8) 1 rotate(8) 2 residual
 * 3 olsen 4 mgs_step 5 axpy 6 device copy (yardstick).  14 doubles. */
int b2d_measure_level1(b2d_ctx* ctx, int slot0, int reps, double* out);

#ifdef __cplusplus
}
#endif
#endif
