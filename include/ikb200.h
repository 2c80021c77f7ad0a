/*
 * ikb200.h -- C-ABI of the B200 device flat assembler (libikb200.so).
 *
 * This is the drop-in boundary for the one hot path of Ikarus that this project
 * accelerates: global FEM assembly of tangent K, residual R and energy E plus the
 * Jacobi-PCG solve that consumes them.  The reference has no FFI of its own; its
 * boundary is the C++ concept Concepts::MatrixFlatAssembler
 * (ikarus/utils/concepts.hh:517-585).  Each entry point below names the reference
 * member function it replaces (paths relative to the reference root).
 * include/ikarus_b200/deviceflatassembler.hh wraps this ABI back into that concept.
 *
 * Conventions
 *  - every function returns IKB_OK (0) or a negative IKB_E* code; ikb_last_error()
 *    gives the message.  No exceptions cross the boundary.
 *  - all pointers are caller-owned HOST buffers unless the name says "device".
 *  - a handle owns one CUDA device + one stream; calls on a handle are serialised by
 *    the caller (the reference assembler is single-threaded and not re-entrant,
 *    ikarus/assembler/simpleassemblers.hh:90-92).
 *  - element-local dof order is node-major/component-minor, i*dim + c
 *    (ikarus/finiteelements/fehelper.hh:145-153).
 *  - matrices use the compressed layout Eigen::SparseMatrix<double> (column-major,
 *    sorted inner indices) has after setFromTriplets
 *    (ikarus/assembler/simpleassemblers.inl:206-251).  The pattern is structurally
 *    symmetric and K is symmetric, so the same arrays are a valid CSR view.
 */
#ifndef IKB200_H
#define IKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 3: ikb_desc.eas_function.  Additions since (no layout change, same version): IKB_MAT_BLATZKO, IKB_MAT_HYPERELASTIC with
 * ikb_set_hyperelastic / ikb_hyperelastic. */
#define IKB_ABI_VERSION 3

typedef struct ikb_handle_s* ikb_handle;

enum { IKB_OK = 0, IKB_EINVAL = -1, IKB_ECUDA = -2, IKB_ESTATE = -3, IKB_ENOTIMPL = -4, IKB_EMATERIAL = -5, IKB_ENCCL = -6 };

/* strain measure of the solid skill: linearElastic (mechanics/linearelastic.hh) or
 * nonLinearElastic (mechanics/nonlinearelastic.hh) */
enum { IKB_STRAIN_LINEAR = 0, IKB_STRAIN_GREEN_LAGRANGE = 1 };
/* Materials::LinearElasticity / StVenantKirchhoff / NeoHooke
 * (mechanics/materials/linearelasticity.hh, svk.hh, hyperelastic/neohooke.hh);
 * IKB_MAT_BLATZKO: Materials::makeBlatzKo(mu) = Hyperelastic<Deviatoric<BlatzKoT>, Volumetric<VF0T>>, the
 * principal-stretch framework (hyperelastic/interface.hh:99-232, deviatoric/interface.hh:77-115,
 * deviatoric/blatzko.hh:60-92, factory.hh:34-39): ikb_desc.mu is its parameter, lambda is ignored; Q1 elements
 * (plain and every EAS form), 3D or planeStrain */
enum { IKB_MAT_LINEAR_ELASTICITY = 0, IKB_MAT_SVK = 1, IKB_MAT_NEOHOOKE = 2, IKB_MAT_BLATZKO = 3, IKB_MAT_HYPERELASTIC = 4 };
/* IKB_MAT_HYPERELASTIC: any Materials::Hyperelastic<Deviatoric<DF>, Volumetric<VF>> of the factories in
 * hyperelastic/factory.hh:12-165, described by ikb_set_hyperelastic below (required before the first assembly) */
/* deviatoric functions: nodeviatoricfunction.hh, blatzko.hh, ogden.hh (PrincipalStretchTags::total | deviatoric),
 * invariantbased.hh (makeMooneyRivlin, makeYeoh, makeInvariantBased), arrudaboyce.hh, gent.hh */
enum { IKB_DEV_NONE = 0, IKB_DEV_BLATZKO = 1, IKB_DEV_OGDEN_TOTAL = 2, IKB_DEV_OGDEN_DEVIATORIC = 3,
       IKB_DEV_INVARIANT_BASED = 4, IKB_DEV_ARRUDA_BOYCE = 5, IKB_DEV_GENT = 6 };
/* DBCOption (assembler/dirichletbcenforcement.hh) */
enum { IKB_DBC_RAW = 0, IKB_DBC_REDUCED = 1, IKB_DBC_FULL = 2 };
/* affordance bits: ScalarAffordance::mechanicalPotentialEnergy, VectorAffordance::forces,
 * MatrixAffordance::stiffness (finiteelements/ferequirements.hh:35-89) */
enum { IKB_SCALAR = 1, IKB_VECTOR = 2, IKB_MATRIX = 4 };

enum { IKB_REDUCE_NONE = 0, IKB_REDUCE_PLANE_STRAIN = 1, IKB_REDUCE_PLANE_STRESS = 2 };
/* EAS::LinearStrain / EAS::GreenLagrangeStrain (by the element's strain tag), EAS::DisplacementGradient,
 * EAS::DisplacementGradientTransposed */
enum { IKB_EAS_STRAIN = 0, IKB_EAS_DISPLACEMENT_GRADIENT = 1, IKB_EAS_DISPLACEMENT_GRADIENT_TRANSPOSED = 2 };

typedef struct ikb_desc {
  int32_t abi_version;  /* IKB_ABI_VERSION */
  int32_t dim;          /* 2 | 3 */
  int32_t order;        /* Lagrange order of the power basis: 1 (Quad4/Hex8) | 2 (Quad9/Hex27) */
  int32_t strain;       /* IKB_STRAIN_* */
  int32_t material;     /* IKB_MAT_* */
  int32_t plane_strain; /* 2D only, IKB_REDUCE_*: Materials::planeStrain(mat) (materials/vanishingstrain.hh) or
                           Materials::planeStress(mat, tol) (materials/vanishingstress.hh) */
  int32_t eas_m;        /* eas<...>(m): 0 | 4,5,7 (2D Q1) | 9,21 (3D Q1)  (easvariants/linearandglstrains.hh);
                           4 (2D) | 9 (3D) with a displacement-gradient eas_function */
  int32_t device;       /* CUDA ordinal, -1 = current device */
  double lambda;        /* Lame's first parameter (physicshelper.hh:53-57) */
  double mu;            /* shear modulus */
  int64_t n_elem;       /* elements owned by this handle */
  int64_t n_dof;        /* global dofs = basis.flat().size() */
  double reduce_tol;    /* planeStress: tolerance of the stress reduction (VanishingStress ctor, default 1e-12) */
  int32_t eas_function; /* IKB_EAS_*: what eas<ES>(m) enhances (the headers under strainenhancements/easfunctions); the
                           displacement-gradient forms take m = dim*dim (H4 / H9, easvariants/displacementgradient.hh)
                           and a nonlinear element */
  int32_t reserved_;
} ikb_desc;

/* ---- lifetime ------------------------------------------------------------------- */
/* replaces makeSparseFlatAssembler / SparseFlatAssembler ctor (assembler/simpleassemblers.hh:174-177) */
int ikb_create(ikb_handle* h, const ikb_desc* desc);
int ikb_destroy(ikb_handle h);
int ikb_last_error(ikb_handle h, char* buf, size_t len);

/* ---- one-time uploads ----------------------------------------------------------- */
/* corner_coords[n_elem][2^dim][dim]: fe.gridElement().geometry().corner(c) (nonlinearelastic.hh:117);
 * elem_dofs[n_elem][nodes*dim]: FEHelper::globalIndices(fe, dofs) (finiteelements/fehelper.hh:194-197).
 * The dofs of one Lagrange node must be FlatInterleaved (dim*node+c) or FlatLexicographic
 * (c*nNodes+node); anything else returns IKB_ENOTIMPL. */
int ikb_upload_mesh(ikb_handle h, const double* corner_coords, const int64_t* elem_dofs);
/* DirichletValues flags (utils/dirichletvalues.hh:73-131); constraintsBelow is derived
 * on the device (FlatAssemblerBase ctor, assembler/interface.hh:51-62). */
int ikb_upload_dirichlet(ikb_handle h, const uint8_t* flags);
/* createOccupationPattern + createLinearDOFsPerElement (+ the Reduced variants)
 * (assembler/simpleassemblers.inl:206-299): builds the sparsity pattern and the
 * deterministic element->CSR gather map on the device. */
int ikb_build_pattern(ikb_handle h);
int ikb_pattern_nnz(ikb_handle h, int dbc, int64_t* rows, int64_t* nnz);
/* outer[rows+1], inner[nnz]: for bit-exact comparison with Eigen's outerIndexPtr/innerIndexPtr */
int ikb_get_pattern(ikb_handle h, int dbc, int64_t* outer, int32_t* inner);
/* constraintsBelow(i) for all i (assembler/interface.hh:124-132) */
int ikb_get_constraints_below(ikb_handle h, int64_t* out);
/* ikb_element_linear_indices: elementLinearIndices_[e] as createLinearDOFsPerElement builds it
 * (column-major over A: for c, for r; simpleassemblers.inl:253-266); Raw pattern only. */
int ikb_element_linear_indices(ikb_handle h, int64_t elem, int64_t* out);

/* ---- per-solve state ------------------------------------------------------------ */
/* Buffer lifetime: ikb_set_solution, ikb_set_solution_range, ikb_set_external_load, ikb_eas_set_alpha and
 * ikb_update_solution enqueue an asynchronous copy from the caller's buffer and return.  With pageable host memory the
 * CUDA runtime has staged the data by then; a PINNED (page-locked) buffer is read when the copy actually runs, so it
 * must stay untouched until ikb_sync(h) -- or any call that returns host data -- has returned. */
/* FERequirements::globalSolution()/parameter() (finiteelements/ferequirements.hh:222-407) */
int ikb_set_solution(ikb_handle h, const double* d);
/* Partial upload: d[0..count) -> resident solution[dof_begin .. dof_begin+count).  An element-partitioned rank only
 * needs the dofs of its owned + ghost node layers. */
int ikb_set_solution_range(ikb_handle h, const double* d, int64_t dof_begin, int64_t count);
int ikb_set_parameter(ikb_handle h, double lambda);
/* Host-sampled volume/Neumann/point loads (mechanics/loads/volume.hh:67-106,
 * loads/traction.hh:70-138): R = F_int - s*fext, E -= s*fext.d with s = lambda when
 * scales_with_lambda else 1. */
int ikb_set_external_load(ikb_handle h, const double* fext, int scales_with_lambda);

/* ---- the hot path --------------------------------------------------------------- */
/* One fused element sweep producing any of K, R, E for one DBC mode; replaces
 * getRawMatrixImpl/getMatrixImpl/getReducedMatrixImpl, get*VectorImpl and getScalarImpl
 * (assembler/simpleassemblers.inl:17-24, 59-204).  Results stay on the device. */
int ikb_assemble(ikb_handle h, unsigned what, int dbc);
/* Marks every cached result stale so that the next ikb_assemble recomputes even if d, lambda and alpha
 * are unchanged (the reference recomputes on every call; the cache is an optimisation of this layer). */
int ikb_invalidate(ikb_handle h);
int ikb_get_vector(ikb_handle h, int dbc, double* out);       /* N or N_red doubles */
int ikb_get_scalar(ikb_handle h, double* energy);
int ikb_get_matrix_values(ikb_handle h, int dbc, double* out); /* nnz doubles, Eigen value order */
/* DenseFlatAssembler (assembler/simpleassemblers.inl:301-375): column-major rows x rows */
int ikb_get_dense_matrix(ikb_handle h, int dbc, double* out);
/* 2-norm of the assembled vector (NewtonRaphson needs ||rx|| on the host,
 * solver/nonlinearsolver/newtonraphson.hh:206) */
int ikb_vector_norm(ikb_handle h, int dbc, double* norm);

/* ---- principal-stretch hyperelastic laws ---------------------------------------- */
typedef struct ikb_hyperelastic {
  int32_t deviatoric;  /* IKB_DEV_* */
  int32_t n;           /* terms of Ogden<n, tag> / InvariantBased<n>, 1..3 (the reference's own tests use up to 3) */
  int32_t volumetric;  /* 0..12: VF0 (none) .. VF12 (volumetric/volumetricfunctions.hh:25-380) */
  int32_t reserved_;
  int32_t pex[3], qex[3]; /* InvariantBased exponents of (W1 - 3), (W2 - 3); not both zero (invariantbased.hh:216-221) */
  double par[3];       /* Ogden: mu_i; InvariantBased: material parameters; BlatzKo: {mu}; ArrudaBoyce: {mu, lambdaM};
                          Gent: {mu, Jm} */
  double ex[3];        /* Ogden: alpha_i */
  double K;            /* Volumetric(matPar, vf): Lame's first parameter for total, the bulk modulus for deviatoric
                          stretches (volumetric/interface.hh:40-52) */
  double beta;         /* VF4, VF7, VF10 */
} ikb_hyperelastic;
/* replaces the Hyperelastic ctor (hyperelastic/interface.hh:74-82) for a handle created with IKB_MAT_HYPERELASTIC */
int ikb_set_hyperelastic(ikb_handle h, const ikb_hyperelastic* law);

/* ---- EAS internal variables ----------------------------------------------------- */
/* EnhancedAssumedStrains::updateStateImpl on CORRECTION_UPDATED
 * (mechanics/enhancedassumedstrains.hh:225-248); correction has N entries (DBCOption::Full). */
int ikb_eas_update(ikb_handle h, const double* correction);
int ikb_eas_get_alpha(ikb_handle h, double* alpha /* [n_elem][m] */);
int ikb_eas_set_alpha(ikb_handle h, const double* alpha);

/* ---- results at local positions ------------------------------------------------- */
/* fe.calculateAt<RT>(req, local) for EVERY element and n_points local positions (xi in [0,1]^dim):
 * NonLinearElastic / LinearElastic / EnhancedAssumedStrains::calculateAtImpl
 * (mechanics/nonlinearelastic.hh:237-271, linearelastic.hh, enhancedassumedstrains.hh:127-187).
 * out[(e*n_points + q)*ncomp + c] in Voigt order (3D [00,11,22,12,02,01], 2D [00,11,01]);
 * ncomp = dim(dim+1)/2, or 6 for the *_FULL types (the underlying 3D law of a plane-strain
 * material).  linearStress* belong to the linear element, the others to the nonlinear one
 * (IKB_ENOTIMPL otherwise, like the reference's supportsResultType). */
enum {
  IKB_RESULT_LINEAR_STRESS = 0,
  IKB_RESULT_PK2_STRESS = 1,
  IKB_RESULT_LINEAR_STRESS_FULL = 2,
  IKB_RESULT_PK2_STRESS_FULL = 3,
  IKB_RESULT_KIRCHHOFF_STRESS = 4,
  IKB_RESULT_CAUCHY_STRESS = 5
};
int ikb_calculate_at(ikb_handle h, int result_type, const double* local, int n_points, double* out);

/* ---- linear solve --------------------------------------------------------------- */
/* Jacobi-preconditioned CG on the assembled matrix of mode dbc (Full or Reduced); the
 * analogue of LinearSolver(SolverTypeTag::si_ConjugateGradient)
 * (solver/linearsolver/linearsolver.cpp:23-24).  rhs == NULL solves K x = -R with the
 * resident residual.  Stops at ||r|| <= rel_tol*||rhs||. */
int ikb_pcg_solve(ikb_handle h, int dbc, const double* rhs, double* x, double rel_tol, int max_it, int* iters,
                  double* rel_res);
/* Steihaug-Toint truncated CG on the assembled matrix of mode dbc: the inner solver of
 * TrustRegion, Eigen::TruncatedConjugateGradient with the Identity or Diagonal preconditioner
 * (linearalgebra/truncatedconjugategradient.hh:68-168, solver/nonlinearsolver/trustregion.hh:529-547),
 * started from x = 0 like TrustRegion::solve does (:262).  rhs == NULL solves H eta = -g with
 * the resident gradient g = R.  x may be NULL: eta stays resident as the correction for
 * ikb_update_solution / ikb_eas_update.  The model terms TrustRegion needs afterwards
 * (|eta|, g.eta, eta.H eta; trustregion.hh:268, 317, 327) are returned so that no vector
 * has to leave the device. */
enum { IKB_PRECOND_IDENTITY = 0, IKB_PRECOND_DIAGONAL = 1 };
enum {
  IKB_TCG_NEGATIVE_CURVATURE = 0,
  IKB_TCG_EXCEEDED_TRUST_REGION = 1,
  IKB_TCG_REACHED_KAPPA_LINEAR = 2,
  IKB_TCG_REACHED_THETA_SUPERLINEAR = 3,
  IKB_TCG_MAXIMUM_INNER_ITERATIONS = 4,
  IKB_TCG_MODEL_INCREASED = 5
}; /* Eigen::TCGStopReason, truncatedconjugategradient.hh:25-33 */
typedef struct ikb_tcg_info {
  /* in (TCGInfo, truncatedconjugategradient.hh:34-50) */
  double delta;       /* trust-region radius */
  double kappa;       /* 0.1 */
  double theta;       /* 1.0 (unused by the reference's stopping rule, kept for layout parity) */
  int64_t mininner;   /* 1 */
  int64_t max_iters;  /* <= 0: 2 n (Eigen::IterativeSolverBase default) */
  double tol;         /* <= 0: machine epsilon (Eigen default) */
  int32_t precond;    /* IKB_PRECOND_* */
  /* out */
  int32_t stop_reason; /* IKB_TCG_* */
  int64_t iterations;
  double rel_error;   /* |r| / |rhs| */
  double eta_norm;    /* |eta|_2 */
  double g_dot_eta;   /* g.eta with g = -rhs */
  double eta_h_eta;   /* eta.(H eta) */
} ikb_tcg_info;
int ikb_tcg_solve(ikb_handle h, int dbc, const double* rhs, double* x, ikb_tcg_info* info);
/* Forces due to inhomogeneous Dirichlet boundary conditions, utils::obtainForcesDueToIDBC
 * (utils/functionhelper.hh:170-185): F = K_raw * d_inc with d_inc = d(d_D)/d(lambda) (N entries, the caller
 * evaluates its boundary functions), then zeroed at constrained dofs (dbc == Full, N entries out) or
 * reduced (otherwise, N_red entries out).  The Raw matrix of the current state is assembled on demand. */
int ikb_idbc_forces(ikb_handle h, int dbc, const double* d_inc, double* out);
/* x += correction on the resident solution (NonlinearSolverFactory update functor,
 * solver/nonlinearsolver/nonlinearsolverfactory.hh:33-56); correction lives on the device
 * (last PCG result) when correction == NULL. dbc selects Reduced->Full expansion. */
int ikb_update_solution(ikb_handle h, int dbc, const double* correction);
int ikb_get_solution(ikb_handle h, double* d);
/* y = K x for the assembled matrix of mode dbc (host in/out); used by tests and by
 * utils::obtainForcesDueToIDBC-style callers (utils/functionhelper.hh:170-185). */
int ikb_spmv(ikb_handle h, int dbc, const double* x, double* y);

/* ---- multi-GPU (element partition, row-block ownership) ------------ */
/* Restricts the handle to the rows of nodes [node_begin, node_end) (a contiguous
 * row block); elements uploaded must be all elements touching an owned node
 * (owner-computes with one ghost layer).  Must precede ikb_build_pattern. */
int ikb_set_row_ownership(ikb_handle h, int64_t node_begin, int64_t node_end);
int ikb_nccl_unique_id(void* id128 /* 128 bytes */);
int ikb_comm_init(ikb_handle h, const void* id128, int rank, int nranks);
/* Pure host helper (no GPU needed): node intervals exchanged between two ranks of a contiguous row-block
 * partition.  mine4/peer4 = {ownBegin, ownEnd, needBegin, needEnd}; out4 = {sendBegin, sendEnd, recvBegin, recvEnd}. */
int ikb_halo_intervals(const int64_t* mine4, const int64_t* peer4, int64_t* out4);
/* exchange the interface node layers of a global-length resident array ("solution" | "correction") */
int ikb_halo_exchange(ikb_handle h, const char* what);
/* Peer-memory transport of the distributed PCG (one process per GPU on one NVLink/NVSwitch node; new, the reference
 * has no distributed code).  After ikb_comm_init every rank exports two CUDA IPC handles (its search-direction
 * vector and its control window, 64 bytes each), the caller all-gathers the 128-byte records of all ranks in rank order
 * and hands them to every rank.  The PCG then stores halo entries and partial dot products directly into the peers'
 * memory from inside its kernels: no NCCL call and no host round trip per iteration.  Optional: without it (or when
 * `ok` comes back 0 because IPC is unavailable) ikb_pcg_solve uses ncclSend/Recv + ncclAllReduce. */
int ikb_comm_ipc_export(ikb_handle h, void* handles128 /* 128 bytes */);
int ikb_comm_ipc_import(ikb_handle h, const void* all_handles /* nranks x 128 bytes */, int* ok);

/* ---- introspection for benchmarks ------------------------------------------------ */
int ikb_stream(ikb_handle h, void** cuda_stream);
int ikb_sync(ikb_handle h);
/* number of kernel launches issued by this handle since creation */
int ikb_launch_count(ikb_handle h, int64_t* n);
/* device pointers of resident arrays: what = "solution" | "residual" | "values" | "correction" */
int ikb_device_ptr(ikb_handle h, const char* what, int dbc, void** ptr);
/* time `reps` back-to-back launches of one phase with CUDA events on the handle's
 * stream: phase = "elements" | "gather" | "spmv" | "dfma_peak" ; returns mean ms */
int ikb_time_phase(ikb_handle h, const char* phase, int dbc, int reps, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* IKB200_H */
