/*
 * edgpu.h -- C ABI of the B200-native Lanczos H x v engine for EDIpack (NORMAL mode).
 *
 * This is the drop-in boundary: plain C symbols, plain pointers and sizes, no torch / C++
 * types.  Each entry point names the reference interface it replaces (paths relative to the
 * EDIpack source tree, v6.1.0).  The Fortran side binds them with `bind(C)` interfaces, see
 * INTEGRATION.md.  All functions returning int return 0 on success and a non-zero code on
 * failure; edgpu_last_error() gives the message (the reference `stop`s with a message, e.g.
 * ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:49,52 -- the Fortran shim turns non-zero into `stop`).
 *
 * State model = the reference's: module-global, one sector "open" at a time
 * (build_Hv_sector_normal ... delete_Hv_sector_normal), not re-entrant, one engine per
 * process (= one MPI rank = one GPU).
 *
 * There is NO CPU fallback: every call fails with an error if no sm_100 device is usable.
 */
#ifndef EDGPU_H
#define EDGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDGPU_MAXORB 5
#define EDGPU_MAXBATH 32
#define EDGPU_UID_BYTES 128

/* bath_type (ED_INPUT_VARS.f90:598) */
enum { EDGPU_BATH_NORMAL = 0, EDGPU_BATH_HYBRID = 1, EDGPU_BATH_REPLICA = 2, EDGPU_BATH_GENERAL = 3 };

/*
 * Everything directMatVec_normal_main reads from module globals
 * (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:23-130, ED_VARS_GLOBAL.f90), as one POD.
 * Spin slot s: 0 = up, 1 = dw; with Nspin=1 the caller stores the same numbers in both
 * (that is what the reference's index `Nspin` does, direct/HxV_dw.f90:14).
 * Orbital / bath indices are 0-based here (1-based in the reference).
 */
typedef struct edgpu_normal_params {
  int32_t Ns;        /* levels per spin, ED_SETUP.f90:118-126 */
  int32_t Norb;
  int32_t Nbath;
  int32_t bath_type; /* EDGPU_BATH_* */
  int32_t hfmode;    /* direct/HxV_local.f90:58 */
  int32_t Nfoo;      /* size(bath_diag,2): Norb, or 1 for hybrid */
  int32_t pad0, pad1;
  double xmu;
  double eloc[2][EDGPU_MAXORB][EDGPU_MAXORB]; /* impHloc(s,s,a,b)+mfHloc(s,s,a,b) */
  double spin_field_z[EDGPU_MAXORB];          /* spin_field(a,3) */
  double exc_field[4];
  double Uloc[EDGPU_MAXORB];                  /* Uloc_internal */
  double Ust[EDGPU_MAXORB][EDGPU_MAXORB];     /* Ust_internal  */
  double Jh[EDGPU_MAXORB][EDGPU_MAXORB];      /* Jh_internal   */
  double Jx[EDGPU_MAXORB][EDGPU_MAXORB];      /* Jx_internal   */
  double Jp[EDGPU_MAXORB][EDGPU_MAXORB];      /* Jp_internal   */
  double diag_hybr[2][EDGPU_MAXORB][EDGPU_MAXBATH];             /* diag_hybr(s,a,k) */
  double bath_diag[2][EDGPU_MAXORB][EDGPU_MAXBATH];             /* bath_diag(s,a|1,k) */
  double hbath[2][EDGPU_MAXORB][EDGPU_MAXORB][EDGPU_MAXBATH];   /* hbath_tmp(s,s,a,b,k) */
  int32_t stride[EDGPU_MAXORB][EDGPU_MAXBATH];                  /* getBathStride(a,k), 1-based */
} edgpu_normal_params;

/*
 * Everything ed_buildH_nonsu2_main reads from module globals
 * (ED_HAMILTONIAN_NONSU2_STORED_HxV.f90:29-190 and ED_NONSU2/stored/*.f90), normal / hybrid bath.
 * Spin index 0 = up, 1 = dw; complex numbers as (re, im).
 */
typedef struct edgpu_nonsu2_params {
  int32_t Ns, Norb, Nbath, bath_type, hfmode, Nfoo, pad0, pad1;
  double xmu;
  double hloc[2][2][EDGPU_MAXORB][EDGPU_MAXORB][2]; /* impHloc(ispin,jspin,iorb,jorb) */
  double spin_field[EDGPU_MAXORB][3];               /* spin_field(iorb, x|y|z) */
  double Uloc[EDGPU_MAXORB];
  double Ust[EDGPU_MAXORB][EDGPU_MAXORB];
  double Jh[EDGPU_MAXORB][EDGPU_MAXORB];
  double Jx[EDGPU_MAXORB][EDGPU_MAXORB];
  double Jp[EDGPU_MAXORB][EDGPU_MAXORB];
  double bath_e[2][EDGPU_MAXORB][EDGPU_MAXBATH];    /* dmft_bath%e(ispin, iorb|1, k) */
  double bath_v[2][EDGPU_MAXORB][EDGPU_MAXBATH];    /* dmft_bath%v(ispin, iorb, k)   */
  double bath_u[2][EDGPU_MAXORB][EDGPU_MAXBATH];    /* dmft_bath%u(ispin, iorb, k): spin-flip hybridisation */
  int32_t stride[EDGPU_MAXORB][EDGPU_MAXBATH];      /* getBathStride(a,k), 1-based */
} edgpu_nonsu2_params;

/*
 * Everything ed_buildH_superc_main reads from module globals
 * (ED_HAMILTONIAN_SUPERC_STORED_HxV.f90:29-260 and ED_SUPERC/stored/*.f90), normal / hybrid bath.
 * With Nspin=1 the caller stores the same numbers in both spin slots (the reference's index
 * `Nspin`, stored/Himp.f90:18).
 */
typedef struct edgpu_superc_params {
  int32_t Ns, Norb, Nbath, bath_type, hfmode, Nfoo, pad0, pad1;
  double xmu;
  double hloc[2][EDGPU_MAXORB][EDGPU_MAXORB][2];         /* impHloc(s,s,a,b)+mfHloc(s,s,a,b) */
  double hloc_anomalous[EDGPU_MAXORB][EDGPU_MAXORB][2];  /* impHloc_anomalous(1,1,a,b) */
  double pair_field[EDGPU_MAXORB];
  double Uloc[EDGPU_MAXORB];
  double Ust[EDGPU_MAXORB][EDGPU_MAXORB];
  double Jh[EDGPU_MAXORB][EDGPU_MAXORB];
  double Jx[EDGPU_MAXORB][EDGPU_MAXORB];
  double Jp[EDGPU_MAXORB][EDGPU_MAXORB];
  double bath_e[2][EDGPU_MAXORB][EDGPU_MAXBATH];   /* dmft_bath%e(ispin, iorb|1, k) */
  double bath_d[EDGPU_MAXORB][EDGPU_MAXBATH];      /* dmft_bath%d(1, iorb|1, k): bath pairing */
  double bath_v[2][EDGPU_MAXORB][EDGPU_MAXBATH];   /* dmft_bath%v(ispin, iorb, k) */
  int32_t stride[EDGPU_MAXORB][EDGPU_MAXBATH];     /* getBathStride(a,k), 1-based */
} edgpu_superc_params;

/*
 * One user two-body operator  U cd_i cd_j c_k c_l  = type coulomb_matrix_element
 * (ED_VARS_GLOBAL.f90:24-31), the entries of the module global `coulomb_sundry` read by
 * direct/HxV_sundry.f90:16-21: element [0] of every operator is the impurity orbital (1-based),
 * element [1] the spin (1 = up, 2 = dw), exactly as the reference stores them.
 */
typedef struct edgpu_sundry_term {
  int32_t cd_i[2], cd_j[2], c_k[2], c_l[2];
  double U;
} edgpu_sundry_term;
#define EDGPU_MAXSUNDRY 1024

/* ---------------- engine / communicator ---------------- */

/* Selects the CUDA device and creates the engine's stream.  Replaces nothing in the
 * reference (it has no device); called once from ed_init_solver (ED_MAIN.f90:90). */
int edgpu_init(int device);
int edgpu_finalize(void); /* ed_finalize_solver, ED_MAIN.f90:224 */

/* Replaces ed_set_MpiComm (ED_VARS_GLOBAL.f90:341-357): rank/size of the dw-split.
 * uid is an NCCL unique id (EDGPU_UID_BYTES) created on rank 0 by edgpu_comm_unique_id and
 * broadcast by the caller (MPI_Bcast in the Fortran host, torch.distributed in bench.py). */
int edgpu_comm_unique_id(void *uid);
int edgpu_comm_init(int rank, int nranks, const void *uid);
int edgpu_comm_rank(void);
int edgpu_comm_size(void);

/* ---------------- sector: build_Hv_sector_normal / delete_Hv_sector_normal ------------ */

/* build_Hv_sector_normal(isector) (ED_HAMILTONIAN_NORMAL.f90:31-204): builds the
 * DEVICE-RESIDENT sector maps (build_sector, ED_SECTOR.f90:165-242), the dw split
 * (:128-142), the per-spin hop tables and diagonal tables, and selects the H x v kernels.
 * The sector is addressed by (nup,ndw) = get_Nup/get_Ndw(isector). */
int edgpu_sector_open_normal(const edgpu_normal_params *p, int nup, int ndw);
/* The module global `coulomb_sundry` (filled by ED_PARSE_UMATRIX from a umatrix file): the
 * user-defined two-body terms applied by direct/HxV_sundry.f90:1-109 (direct_mpi twin) as
 * Hv(j) += U sg1 sg2 sg3 sg4 vin(i), operators applied right to left as c_l, cd_j, c_k, cd_i on
 * the row state.  Engine-global like the reference's: applies to every NORMAL sector opened
 * AFTER the call; nterms = 0 deallocates it.  A term that changes the total spin is refused
 * (the reference STOPs, HxV_sundry.f90:35). */
int edgpu_set_coulomb_sundry(int nterms, const edgpu_sundry_term *terms);
/* Phonon inputs Nph, w0_ph, A_ph, g_ph(Norb,Norb) (ED_INPUT_VARS.f90:184-198), read by
 * direct/HxV_ph.f90:1-6 and direct/HxV_eph.f90:1-81: DimPh = Nph+1 (ED_SETUP.f90:137), every
 * vector of a NORMAL sector opened afterwards holds DimPh consecutive electronic chunks,
 * i = i_el + (iph-1)*DimUp*mpiQdw (direct_mpi/HxV_eph.f90:3-4), and H gains
 * w0 b^+b + sum_ab g(a,b) c^+_a c_b (b + b^+) [+ A_ph (b + b^+) as the STORED path has it,
 * stored/H_ph.f90:6-17 -- the direct path has no A_ph term; pass 0 to reproduce it].
 * g_ph is row-major g_ph[a*Norb + b]; Nph = 0 switches phonons off. */
int edgpu_set_phonons(int Nph, double w0_ph, double A_ph, const double *g_ph, int Norb);
/* ed_total_ud = F: build_Hv_sector_normal for an orbital-resolved sector (Nups(1:Norb), Ndws(1:Norb))
 * (Ns_Ud = Norb factors of Ns_Orb = 1+Nbath levels, ED_SETUP.f90:128-135; maps build_sector
 * ED_SECTOR.f90:217-242; state index = mixed radix over [DimUps, DimDws], :1691-1702) and the
 * Hamiltonian of directMatVec_normal_orbs / ed_buildh_normal_orbs
 * (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:133-226, direct/Orbs/HxV_local.f90, HxV_up.f90, HxV_dw.f90),
 * generated on the device as a real stored H.  Afterwards the sector behaves like one opened with
 * edgpu_csr_open_d (edgpu_hxv_d, Lanczos drivers, edgpu_eigh; edgpu_csr_get downloads it).  Needs
 * bath_type = normal and no inter-orbital term (the reference stops otherwise, ED_SETUP.f90:124,
 * ED_PARSE_UMATRIX.f90:156-163).  With nranks>1 the rows are split along the last factor. */
int edgpu_sector_open_normal_orbs(const edgpu_normal_params *p, const int32_t *nups, const int32_t *ndws);
int edgpu_sector_close(void);           /* delete_Hv_sector_normal, :212-279 */
int64_t edgpu_sector_vecdim(void);      /* vecDim_Hv_sector_normal, :286-313 (local chunk) */
int64_t edgpu_sector_dim(void);         /* getDim(isector) */
int edgpu_sector_dims(int64_t *DimUp, int64_t *DimDw, int64_t *qdw, int64_t *dw_start);
/* How the Hdw term of the open sector crosses ranks (replaces vector_transpose_MPI,
 * ED_HAMILTONIAN_NORMAL_COMMON.f90:66-178): mode 0 = single rank, 1 = halo (the owners store the
 * remote dw columns this rank's hops read into its halo buffer over NVLink), 2 = chunk-pipelined
 * peer-memory transposes, 3 = NCCL grouped send/recv transposes.  halo_cols = columns received,
 * send_cols = columns sent per product (each DimUp doubles, padded to a multiple of 16). */
int edgpu_sector_comm_info(int *mode, int64_t *halo_cols, int64_t *send_cols, int *nchunks);
/* The halo exchange plan as a pure host function (no GPU needed; what edgpu_sector_open_normal
 * derives on every rank from the all-gathered need maps).  need[r * dim_dw + d] != 0 when the dw
 * hops of rank r's chunk read column d of another rank (chunks = the dw split of
 * ED_HAMILTONIAN_NORMAL.f90:128-142).  Out: halo_cols[r] = columns in rank r's halo (its needed
 * columns in ascending order); send_triples = (local column of `rank`, destination rank, slot in
 * the destination's halo) for every column `rank` must push, at most send_cap of them written,
 * *nsend = their number. */
int edgpu_halo_plan(int64_t dim_dw, int nranks, int rank, const unsigned char *need, int64_t *halo_cols,
                    int32_t *send_triples, int64_t send_cap, int64_t *nsend);

/* Parity hooks: download the device-resident structures for bit-exact comparison with the
 * oracle (sector maps ED_SECTOR.f90:217-242; hop tables = spH0ups(1)/spH0dws(1) content
 * of stored/H_up.f90, H_dw.f90, one row per source state, targets 1-based). */
int edgpu_sector_get_map(int spin, int32_t *map);
int64_t edgpu_sector_hop_count(int spin);                      /* total entries */
int edgpu_sector_get_hops(int spin, int64_t *rowptr, int32_t *target, double *value);

/* ---------------- stored-H sector (ED_SPARSE_H=T) ---------------- */

/* The host's stored Hamiltonian as CSR.  Replaces the sparse_matrix_csr objects spH0 /
 * spH0d+spH0ups+spH0dws+spH0nd (ED_SPARSE_MATRIX.f90:16-41, filled by ed_buildh_*_main,
 * ..._STORED_HxV.f90) and their products spMatVec[_mpi]_{normal,superc,nonsu2}_main
 * (ED_HAMILTONIAN_NORMAL_STORED_HxV.f90:517-929, ..._NONSU2_STORED_HxV.f90:194-265,
 * ..._SUPERC_STORED_HxV.f90:312-432): any ed_mode, real (_d) or complex (_z, values as
 * interleaved re,im).  rowptr[nloc+1] (0-based offsets), cols 1-based GLOBAL column indices as
 * in the reference's row%cols, unsorted and with duplicates allowed (they add up, like
 * sp_insert_element).  Rows are this rank's rows [row_offset, row_offset+nloc) of the flat
 * row split MpiQ = Dim/P, remainder to the last rank (ED_HAMILTONIAN_NONSU2.f90:72-79);
 * with nranks>1 the input vector is all-gathered on every product like the reference does.
 * While a stored-H sector is open, edgpu_hxv_d / edgpu_hxv_z and the Lanczos drivers act on
 * it; vectors are plain arrays of nloc reals / complex numbers. */
int edgpu_csr_open_d(int64_t nloc, int64_t nglobal, int64_t row_offset, const int64_t *rowptr,
                     const int32_t *cols, const double *vals);
int edgpu_csr_open_z(int64_t nloc, int64_t nglobal, int64_t row_offset, const int64_t *rowptr,
                     const int32_t *cols, const double *vals_re_im);

/* ED_SPARSE_H (ED_INPUT_VARS.f90:664) for the packed-state modes (nonsu2 / superc), read by the
 * next edgpu_sector_open_nonsu2 / _superc: 1 (default) = the complex spH0 is generated and stored
 * on the device (ed_buildH_nonsu2_main); 0 = nothing is stored and every product re-enumerates the
 * matrix elements (directMatVec_nonsu2_main / directMatVec_superc_main,
 * ED_HAMILTONIAN_NONSU2_DIRECT_HxV.f90:22-252, ED_HAMILTONIAN_SUPERC_DIRECT_HxV.f90:22-311).  The flag
 * also selects directMatVec_normal_orbs (ED_HAMILTONIAN_NORMAL_DIRECT_HxV.f90:134-227) for
 * edgpu_sector_open_normal_orbs.  ed_total_ud=T NORMAL sectors are always direct (hop tables, no
 * stored matrix). */
int edgpu_set_sparse_h(int flag);
/* build_Hv_sector_nonsu2(isector) + ed_buildH_nonsu2_main (ED_HAMILTONIAN_NONSU2.f90:31-130,
 * ED_HAMILTONIAN_NONSU2_STORED_HxV.f90:29-190) entirely on the device: the sector map
 * H(1)%map(DimEl) of packed states m = iup + idw*2^Ns with Ntot electrons (build_sector,
 * ED_SECTOR.f90:351-368), this rank's rows of the flat split (:72-79) and the complex stored
 * Hamiltonian spH0 (element generators ED_NONSU2/stored/Himp.f90, Hint.f90, Hbath.f90,
 * Himp_bath.f90) are generated by kernels; afterwards the sector behaves like one opened with
 * edgpu_csr_open_z (edgpu_hxv_z, Lanczos drivers, edgpu_eigh).  Normal and hybrid baths.
 * Parity hooks: edgpu_sector_get_map(0, map) returns the DimEl packed states;
 * edgpu_csr_nnz / edgpu_csr_get download the device CSR (rowptr 0-based offsets, cols 1-based
 * global, vals (re,im) pairs for complex sectors), duplicates within a row add up. */
int edgpu_sector_open_nonsu2(const edgpu_nonsu2_params *p, int ntot);
/* bath_type = replica / general for the two device-built modes: the module array
 * Hbath_tmp(:,:,:,:,ibath) = build_Hreplica / build_Hgeneral(dmft_bath%item(ibath)%lambda)
 * (ED_HAMILTONIAN_NONSU2_STORED_HxV.f90:82-105, ED_HAMILTONIAN_SUPERC_STORED_HxV.f90:80-112) read by
 * ED_NONSU2/stored/Hbath.f90:29-133 (diagonal, same-spin and spin-flip bath hops) and
 * ED_SUPERC/stored/Hbath.f90:29-92,135-177 (Nambu blocks: (1,1) particles, (2,2) holes, (1,2)/(2,1)
 * pairing), as complex numbers hbath[is][js][a][b][k] (re,im), row-major
 * [2][2][Norb][Norb][Nbath].  Engine-global: read by every nonsu2 / superc sector opened afterwards
 * with bath_type replica / general (their params carry diag_hybr in bath_v; bath_e, bath_u, bath_d
 * are not used).  NULL clears it. */
int edgpu_set_hbath_packed(const double *hbath_re_im, int Norb, int Nbath);
/* build_Hv_sector_superc(isector) + ed_buildH_superc_main (ED_HAMILTONIAN_SUPERC.f90:31-140,
 * ED_HAMILTONIAN_SUPERC_STORED_HxV.f90:29-260) on the device, sector Sz = Nup - Ndw
 * (build_sector, ED_SECTOR.f90:244-281): same contract as edgpu_sector_open_nonsu2, with the
 * anomalous local terms (impHloc_anomalous, pair_field) and the bath pairing amplitudes d. */
int edgpu_sector_open_superc(const edgpu_superc_params *p, int sz);
int64_t edgpu_csr_nnz(void);
int edgpu_csr_get(int64_t *rowptr, int32_t *cols, double *vals);

/* ---------------- H x v ---------------- */

/* Signature-compatible with the abstract interface dd_sparse_HxV(Nloc,v,Hv)
 * (ED_VARS_GLOBAL.f90:111-120) so that `spHtimesV_p => edgpu_hxv_d` works: HOST arrays of
 * the local chunk length, Hv fully overwritten, v untouched.  Errors are latched and
 * reported by edgpu_last_error()/edgpu_status(). */
void edgpu_hxv_d(const int32_t *Nloc, const double *v, double *Hv);
/* cc_sparse_HxV(Nloc,v,Hv) (ED_VARS_GLOBAL.f90:121-132), complex(8) as interleaved re,im:
 * `spHtimesV_cc => edgpu_hxv_z` for a complex stored-H sector. */
void edgpu_hxv_z(const int32_t *Nloc, const double *v_re_im, double *Hv_re_im);
int edgpu_status(void);
/* Page-locks a HOST array of the caller (cudaHostRegister) so that the copies inside edgpu_hxv_d /
 * edgpu_hxv_z / the drivers' host hand-offs run at PCIe speed instead of through the driver's
 * pageable staging.  Meant for the Lanczos work vectors the Fortran side allocates once per sector
 * (ED_DIAG_NORMAL.f90:200-213: `allocate(eig_basis_tmp(...))`).  Unregister before deallocate. */
int edgpu_host_register(void *ptr, int64_t bytes);
int edgpu_host_unregister(void *ptr);

/* Device-resident variant on the engine's internal (padded) layout; d_v/d_Hv are device
 * pointers obtained from edgpu_vec_* below. */
int edgpu_hxv_dev(const double *d_v, double *d_Hv);

/* Internal layout helpers: vectors live on the device as qdw columns of length DimUp with a
 * padded leading dimension.  upload/download convert from/to the reference's contiguous
 * chunk layout i = iup + (idw-1)*DimUp (ED_SECTOR.f90:1681). */
int64_t edgpu_vec_padded_len(void);
int edgpu_vec_upload(double *d_dst, const double *h_src);
int edgpu_vec_download(double *h_dst, const double *d_src);

/* ---------------- Lanczos drivers (SciFortran SF_SP_LINALG replacements) -------------- */

/* sp_lanc_eigh([MpiComm,]MatVec,egs,vect,Nitermax,threshold=) as called at
 * ED_DIAG_NORMAL.f90:206-213: two-pass plain Lanczos ground state with all vectors in HBM.
 * vec_host: in  -- start vector (local chunk) if use_start != 0, else a seeded random start;
 *           out -- normalised ground-state chunk (may be NULL: the vector then only stays
 *                  resident on the device as the "current state", see edgpu_state_*). */
int edgpu_lanczos_gs(int nitermax, double threshold, int ncheck, int use_start, uint64_t seed,
                     double *egs, double *vec_host, int *niter);

/* Diagnostics of the last edgpu_lanczos_gs: Lanczos vectors that were kept in HBM during pass 1
 * (all of them when device memory allows: pass 2 is then a linear combination instead of a
 * second run of the recurrence) and the number of H x v products spent in both passes. */
int edgpu_lanczos_last_info(int *nstored, int *nhxv);
/* The buffers holding those vectors are pooled across solves and sectors (allocating HBM costs
 * more than a second run of the recurrence saves on a single solve).  The library gives the pool
 * back by itself when one of its own allocations runs short; call this to return it earlier. */
int edgpu_release_cache(void);

/* sp_lanc_tridiag([MpiComm,]MatVec,vin,alanc,blanc) as called from
 * tridiag_Hv_sector_normal (ED_HAMILTONIAN_NORMAL.f90:321-369).  seed_host = local chunk of
 * the (un-normalised) start vector, or NULL to use the device-resident seed produced by
 * edgpu_apply_op.  Outputs: alanc[nlanc], blanc[nlanc] (blanc[0] unused like blanc(1)),
 * *nused iterations actually done (early exit when beta < threshold), *norm2 = <seed|seed>. */
int edgpu_lanczos_tridiag(const double *seed_host, int nlanc, double threshold, double *alanc,
                          double *blanc, int *nused, double *norm2);

/* sp_eigh([MpiComm,]MatVec,eval(Neigen),evec(Nloc,Neigen),Nblock,Nitermax,tol=) as called at
 * ED_DIAG_NORMAL.f90:179-192 (LANC_METHOD=arpack, the reference's default; ED_DIAG_NONSU2.f90:179,
 * ED_DIAG_SUPERC.f90:161 for complex sectors): the `neigen` lowest eigenpairs of the open sector.
 * ARPACK (dsaupd/dseupd 'SA', znaupd 'SR') is replaced by a thick-restart Lanczos process with a
 * fully re-orthogonalised basis of `nblock` (= ARPACK's ncv, <= 128) vectors resident in HBM;
 * `nitermax` bounds the restarts, `tol` is ARPACK's tolerance on the Ritz estimates
 * (|beta y_last| <= tol*max(eps^(2/3),|theta|), floored at machine precision).
 * evals[neigen] ascending; evecs_host = NULL or neigen consecutive local chunks (real: Nloc
 * doubles each, complex: Nloc (re,im) pairs), orthonormal.  The eigenvectors also stay on the
 * device until the sector is closed: edgpu_eigh_state_store(k, slot) keeps eigenvector k
 * (0-based) as state `slot` (es_add_state, ED_DIAG_NORMAL.f90:262-278). */
int edgpu_eigh(int neigen, int nblock, int nitermax, double tol, uint64_t seed, double *evals,
               double *evecs_host, int *nconv, int *nmatvec);
int edgpu_eigh_state_store(int k, int slot);

/* ---------------- state hand-off (ED_EIGENSPACE es_add_state / es_return_dvec) --------- */

/* Keeps the last edgpu_lanczos_gs eigenvector on the device as state `slot` together with
 * its sector, so that Green's-function seeds never visit the host. */
int edgpu_state_store(int slot);
int edgpu_state_free(int slot);
/* es_return_dvector / es_return_cvector (ED_EIGENSPACE.f90:620-793): this rank's chunk of the
 * stored state in the reference's layout (real: vecDim doubles; complex sectors: vecDim (re,im)
 * pairs).  The state's own sector must be open. */
int edgpu_state_download(int slot, double *vec_host);
/* apply_op_C (op=-1) / apply_op_CDG (op=+1) of ED_SECTOR.f90:465 / :654 on the stored state:
 * builds c_{iorb,spin}|state> (iorb 0-based, spin 0 up / 1 dw) directly in the layout of the
 * CURRENTLY OPEN sector, which must be the target sector getCsector/getCDGsector. */
int edgpu_apply_op(int slot, int op, int iorb, int spin);
/* apply_COps (ED_SECTOR.f90, nonsu2 / superc branches; used by ED_GF_NONSU2.f90:159-301,
 * ED_GF_SUPERC.f90 and the order-parameter observables ED_OBSERVABLES_SUPERC.f90:204-232): the seed
 * sum_k coef_k O_k |state>, O_k = c^+ (op=+1) / c (op=-1) on (iorb_k, spin_k), 1 <= nops <= 4, built on
 * the device in the layout of the open device-built sector, which must be the common target sector
 * of all operators (Ntot +- 1 for nonsu2; Sz +- 1 for superc).  coef = nops (re,im) pairs.
 * edgpu_apply_op works on these sectors too (nops = 1, coef = 1).  edgpu_seed_norm2 returns
 * <seed|seed> (all-reduced) of the device-resident seed. */
int edgpu_apply_ops_packed(int slot, int nops, const double *coef_re_im, const int *op, const int *iorb,
                           const int *spin);
/* apply_Cops in NORMAL mode (ED_SECTOR.f90 apply_Cops; lanc_build_gf_normal_mix,
 * ED_GF_NORMAL.f90:211,227: vvinit = apply_Cops(v_state,[1,1],[+-1,+-1],[iorb,jorb],[ispin,ispin],...)):
 * the device-resident seed sum_k coef[k] O_k |state>, every O_k = c^+ (op=+1) or c (op=-1) of the
 * same spin on impurity orbital iorb[k] (0-based), in the layout of the OPEN sector, which must be
 * the common target sector.  Needed for the off-diagonal impurity Green's function G_ab. */
int edgpu_apply_ops_normal(int slot, int nops, const double *coef, int op, const int *iorb, int spin);
int edgpu_seed_norm2(double *norm2);
/* Twin states (ED_TWIN=T): es_return_dvector / es_return_cvector for a state with itwin set
 * (ED_EIGENSPACE.f90:640-660, 723-793) = the eigenvector of the twin sector re-ordered by
 * twin_sector_order (ED_SECTOR.f90:1747-1776: Fock states flipped by flip_state_normal / _other
 * :1787-1817, sort_array returns the sorting permutation :1866-1879).  State `dst_slot` := state
 * `src_slot` re-expressed in its twin sector, which must be the OPEN sector: normal
 * (nup,ndw) -> (ndw,nup) [the DimUp x DimDw matrix transposed], superc Sz -> -Sz [up and dw halves of
 * the packed state exchanged], nonsu2 Ntot -> 2Ns-Ntot [every bit complemented].  Stays on the
 * device (the reference gathers to the master and permutes there). */
int edgpu_state_twin(int src_slot, int dst_slot);
/* dens(a), docc(a) of ED_OBSERVABLES_NORMAL.f90:150-215 (ED_OBSERVABLES_NONSU2 / _SUPERC :150-165
 * for packed-state sectors) for the stored state (weight 1). */
int edgpu_state_observables(int slot, double *dens, double *docc);

/* ---------------- diagnostics ---------------- */
const char *edgpu_last_error(void);
/* number of kernels launched by this library since the last reset (bench "gpu_launches") */
int64_t edgpu_launch_count(int reset);
/* elapsed device time (ms) of the H x v kernels of the last edgpu_hxv_* call, per stage:
 * out[0]=up+diag kernel, out[1]=dw kernel, out[2]=non-local kernel, out[3]=transposes */
int edgpu_last_hxv_stage_ms(float *out4);
/* select kernel variant: 0 = auto, 1 = generic gather kernels, 2 = shared-memory tiled */
int edgpu_set_kernel_variant(int variant);
/* the cudaStream_t every kernel of this library is launched on (so that callers can put their
 * own CUDA events on it) */
void *edgpu_stream(void);
/* Per-stage device timing over a run of edgpu_hxv_dev calls without host synchronisation:
 * begin arms a ring of CUDA events for up to max_steps calls; end synchronises and returns the
 * SUMS (ms) over the recorded calls: ms[0]=up+diag, ms[1]=dw (incl. transposes when nranks>1),
 * ms[2]=non-local, and the number of calls recorded. */
int edgpu_profile_begin(int max_steps);
int edgpu_profile_end(float *ms3, int *nsteps);

#ifdef __cplusplus
}
#endif
#endif
