/* piclas_gpu.h — C ABI of the B200 particle step for PICLas' PIC-Poisson (HDG) solver.
 *
 * The reference has no plugin/FFI boundary around its particle path: the "interface" is a set of
 * argument-less Fortran subroutines working on module-global arrays, called from one time-step routine
 * (reference src/timedisc/timedisc_TimeStepPoissonByBorisLeapfrog.f90:93-279 and the Leapfrog sibling
 * src/timedisc/timedisc_TimeStepPoisson.f90:90-284).  This header DEFINES the boundary a Fortran host binds
 * with ISO_C_BINDING (the in-tree precedent for the style is src/output/output.f90:33-48); the Fortran side
 * is in piclas_b200/fortran/mod_particle_gpu.f90 and INTEGRATION.md.
 *
 * Conventions (so that Fortran arrays can be passed as they are):
 *   - every pointer is to host memory in Fortran column-major order; the C index order given in the
 *     comments is the Fortran order reversed, e.g. PartState(1:6,1:n) == double[n][6];
 *   - element/side/node ids are 1-based exactly as in the Fortran tables, except ElemSideNodeID, which
 *     the reference itself stores 0-based ("+1 at use", particle_mesh_tools.f90:1885-1895);
 *   - INTEGER == int32_t, REAL == double, LOGICAL == int32_t (0 = .FALSE.);
 *   - every function returns 0 on success, nonzero on error (the host then calls Abort(__STAMP__,
 *     piclas_gpu_last_error()) as the replaced code does, globals/globals.f90:322-397); nothing here
 *     ever calls exit();
 *   - the host keeps ownership of everything it passes; the library copies at the call;
 *   - one MPI rank == one GPU; calls arrive from the rank's main thread; every call is synchronous
 *     at return.
 */
#ifndef PICLAS_GPU_H
#define PICLAS_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TrackingMethod, src/piclas.h:357-359 */
#define PGPU_REFMAPPING    1
#define PGPU_TRACING       2   /* not supported */
#define PGPU_TRIATRACKING  3

/* DepositionType as the string->integer map of pic_depo_method.f90:39-44 */
#define PGPU_DEPO_CVW          0   /* cell_volweight               (not supported) */
#define PGPU_DEPO_SF           1   /* shape_function                */
#define PGPU_DEPO_SF_CC        2   /* shape_function_cc             */
#define PGPU_DEPO_SF_ADAPTIVE  3   /* shape_function_adaptive       */
#define PGPU_DEPO_CVWM         6   /* cell_volweight_mean           */

/* PP_TimeDiscMethod values of the two time-step routines that own this path */
#define PGPU_TIMEDISC_LEAPFROG        509  /* timedisc_TimeStepPoisson.f90            */
#define PGPU_TIMEDISC_BORIS_LEAPFROG  508  /* timedisc_TimeStepPoissonByBorisLeapfrog */

/* PartBound%TargetBoundCond, particle_boundary_condition.f90:167-214 */
#define PGPU_BC_OPEN        1
#define PGPU_BC_REFLECTIVE  2   /* specular wall at rest (MomentumACC = 0, WallVelo = 0) */
#define PGPU_BC_PERIODIC    3

/* ---- mesh + basis tables (built once by InitParticleMesh / InitializeDeposition) ----------------------
 * Arrays are GLOBAL (all nGlobalElems elements, as the reference's MPI-3 shared windows *_Shared,
 * particle_mesh_vars.f90:109-205) unless marked LOCAL (offsetElem+1 .. offsetElem+nElems). */
typedef struct pgpu_mesh {
  int32_t nGlobalElems, nSides, nNonUniqueNodes, nUniqueGlobalNodes;
  int32_t NGeo;              /* geometry degree; this build requires NGeo == 1                          */
  int32_t N;                 /* solution degree, uniform: N_DG_Mapping(2,:) all equal (dg_vars.f90:38)   */
  int32_t offsetElem;        /* first global element of this rank minus 1 (mesh_vars offsetElem)         */
  int32_t nElems;            /* local element count                                                      */
  int32_t elemInfoSize;      /* leading dimension of ElemInfo (ELEMINFOSIZE = 8 with MPI, piclas.h:143)  */
  int32_t sideInfoSize;      /* leading dimension of SideInfo (SIDEINFOSIZE = 8, piclas.h:166)           */
  const int32_t *ElemInfo;        /* [nGlobalElems][elemInfoSize]; cols piclas.h:149-158, 7 = ELEM_RANK  */
  const int32_t *SideInfo;        /* [nSides][sideInfoSize]; cols piclas.h:167-175                       */
  const double  *NodeCoords;      /* [nNonUniqueNodes][3]                                                */
  const int32_t *NodeInfo;        /* [nNonUniqueNodes] unique node id                                    */
  const int32_t *ElemNodeID;      /* [nGlobalElems][8]  non-unique node ids in CGNS corner order         */
  const int32_t *ElemSideNodeID;  /* [nGlobalElems][6][4]  0-based non-unique node ids                   */
  const int32_t *ConcaveElemSide; /* [nGlobalElems][6]  LOGICAL                                          */
  const double  *XCL_NGeo;        /* [nGlobalElems][NGeo+1][NGeo+1][NGeo+1][3]                           */
  const double  *dXCL_NGeo;       /* [nGlobalElems][NGeo+1][NGeo+1][NGeo+1][3][3]  (Fortran (dd,nn,i,j,k)) */
  const double  *XiCL_NGeo;       /* [NGeo+1]                                                            */
  const double  *wBaryCL_NGeo;    /* [NGeo+1]                                                            */
  const double  *ElemBaryNGeo;    /* [nGlobalElems][3]                                                   */
  const double  *ElemRadius2NGeo; /* [nGlobalElems]                                                      */
  const double  *XiEtaZetaBasis;  /* [nGlobalElems][6][3]                                                */
  const double  *slenXiEtaZetaBasis; /* [nGlobalElems][6]                                                */
  const double  *xGP, *wGP, *wBary;  /* [N+1] N_Inter(N)%xGP / wGP / wBary (interpolation_vars.f90:25-51)*/
  const double  *Elem_xGP;        /* [nGlobalElems][N+1][N+1][N+1][3]  (k,j,i order)                     */
  const double  *ElemsJ;          /* [nGlobalElems][N+1][N+1][N+1]                                       */
  int32_t nBCs;
  const int32_t *bc_kind;         /* [nBCs] PartBound%TargetBoundCond(PartBound%MapToPartBC(BCID))       */
  const int32_t *bc_alpha;        /* [nBCs] BoundaryType(BCID,BC_ALPHA): signed periodic-vector id       */
  int32_t nPeriodicVectors;
  const double  *PeriodicVectors; /* [nPeriodicVectors][3]  GEO%PeriodicVectors                          */
  /* cell_volweight_mean (pic_depo_vars.f90:118-154) */
  const int32_t *Periodic_nNodes;     /* [nUniqueGlobalNodes]                                            */
  const int32_t *Periodic_offsetNode; /* [nUniqueGlobalNodes]                                            */
  const int32_t *Periodic_Nodes;      /* [nPeriodicNodesTotal] unique node ids                           */
  int32_t nPeriodicNodesTotal;
  const double  *NodeVolume;      /* [nUniqueGlobalNodes]                                                */
  /* fast-init background mesh FIBGM (particle_mesh_vars.f90:70-73, GEO%FIBGM*) — RefMapping / shape function */
  double  FIBGMdeltas[3];
  double  xyzminglob[3], xyzmaxglob[3];   /* GEO%xminglob .. GEO%zmaxglob */
  int32_t FIBGMmin[3], FIBGMmax[3];       /* GEO%FIBGMimin..kmax */
  const int32_t *FIBGM_nElems;    /* [kmax-kmin+1][jmax-jmin+1][imax-imin+1]                             */
  const int32_t *FIBGM_offsetElem;/* same shape                                                          */
  const int32_t *FIBGM_Element;   /* [nFIBGMElemsTotal] global element ids                               */
  int32_t nFIBGMElemsTotal;
  /* RefMapping only */
  const double  *ElemEpsOneCell;  /* [nGlobalElems]                                                      */
  const int32_t *ElemToBCSides;   /* [nGlobalElems][2]  (ELEM_NBR_BCSIDES, ELEM_FIRST_BCSIDE)            */
  const double  *SideBCMetrics;   /* [nBCSidesTotal][7]  REAL-typed (particle_mesh_vars.f90:83-84)       */
  int32_t nBCSidesTotal;
  const int32_t *SideType;        /* [nSides] PLANAR_RECT=0 ...                                          */
  const double  *SideNormVec;     /* [nSides][3]                                                         */
  const double  *SideDistance;    /* [nSides]                                                            */
  const double  *BaseVectors0, *BaseVectors1, *BaseVectors2; /* [nSides][3]                              */
  const double  *BaseVectorsScale;/* [nSides]                                                            */
  /* shape function */
  const double  *SFElemr2;        /* [nGlobalElems][2] adaptive radius (r, r^2) or NULL                  */
  const double  *ElemRadiusNGeo;  /* [nGlobalElems] (particle_mesh_build.f90:300)                        */
  const int32_t *ElemToBGM;       /* [nGlobalElems][6] FIBGM cell box imin,imax,jmin,jmax,kmin,kmax (particle_bgm.f90:517-522) */
  const double  *BaseVectors3;    /* [nSides][3] bilinear term of the side (particle_mesh_build.f90:1923); needed when a BC side is
                                     PLANAR_NONRECT or BILINEAR (ComputeBiLinearIntersection), may be NULL otherwise           */
} pgpu_mesh_t;

/* ---- already-parsed run-time parameters (parameter.ini keys in the comments) -------------------------- */
typedef struct pgpu_params {
  int32_t TrackingMethod;       /* TrackingMethod                                  particle_mesh.f90:62   */
  int32_t RefMappingGuess;      /* RefMappingGuess (1..4)                          particle_mesh.f90:352  */
  double  RefMappingEps;        /* RefMappingEps  (default 1e-4)                   particle_mesh.f90:366  */
  int32_t CartesianPeriodic;    /* CartesianPeriodic (must be 0)                   particle_mesh.f90:97   */
  int32_t TimeDiscMethod;       /* PGPU_TIMEDISC_*  (compile-time PP_TimeDiscMethod in the reference)    */
  int32_t DoInterpolation;      /* PIC-DoInterpolation                             pic_interpolation.f90  */
  int32_t DoDeposition;         /* PIC-DoDeposition                                                      */
  int32_t DepositionType;       /* PIC-Deposition-Type as PGPU_DEPO_*                                    */
  double  externalField[6];     /* PIC-externalField                               pic_interpolation.f90:67 */
  double  c2_inv;               /* 1/c^2 (globals_vars.f90:87-90; variable with READIN_CONSTANTS)        */
  int32_t nSpecies;
  const double *ChargeIC;       /* [nSpecies] Part-Species$-ChargeIC                                     */
  const double *MassIC;         /* [nSpecies] Part-Species$-MassIC                                       */
  const double *MacroParticleFactor; /* [nSpecies] Part-Species$-MacroParticleFactor                     */
  /* shape function (pic_depo.f90:55-75, pic_depo_shapefunction_tools.f90:1131-1294) */
  double  r_sf;                 /* PIC-shapefunction-radius                                              */
  int32_t alpha_sf;             /* PIC-shapefunction-alpha                                               */
  int32_t dim_sf;               /* PIC-shapefunction-dimension                                           */
  int32_t dim_sf_dir;           /* PIC-shapefunction-direction                                           */
  int32_t sfDepo3D;             /* PIC-shapefunction-3D-deposition                                       */
  double  w_sf;                 /* normalisation weight computed by InitShapeFunctionDimensionalty       */
  double  dimFactorSF;
  /* device layer */
  int32_t device;               /* CUDA device ordinal                                                   */
  int32_t myRank, nRanks;       /* rank in MPI_COMM_PICLAS; partition given by ElemInfo(ELEM_RANK,:)     */
  int64_t maxParticleNumber;    /* PDM%maxParticleNumber (capacity of the device SoA)                    */
  int32_t carryParticleIDs;     /* 1: carry a 64-bit id per particle through sort/migration (tests)      */
  int32_t arithmetic;           /* 0: reference operation order everywhere; 1: restructured (<=1e-12)    */
  /* options of the reference whose non-default values change the sources / the push and are NOT implemented: piclas_gpu_init
   * fails for any value but 0 ("abort, never ignore") */
  int32_t PartLorentzType;      /* Part-LorentzType: 0 = non-relativistic (default)  particle_rhs.f90:90-123 */
  int32_t NoDirichletDeposition;/* .NOT. PIC-DoDirichletDeposition (NullifyNodeSourceDirichletSides) pic_depo_method.f90:692-697 */
  int32_t DoDielectricSurfaceCharge; /* NodeSourceExt / NodeSourceExtTmp                pic_depo.f90:281-284 */
} pgpu_params_t;

/* after InitParticleMesh + InitializeDeposition (piclaslib.f90:177) */
int piclas_gpu_init(const pgpu_mesh_t *mesh, const pgpu_params_t *params);
int piclas_gpu_finalize(void);
const char *piclas_gpu_last_error(void);

/* host AoS -> device SoA; after ParticleRestart / initial ParticleInserting and after any host-side
 * insertion.  append=0 replaces the device population, append=1 adds to it.
 * PartState[n][6], PartSpecies[n] (1-based), GlobalElemID[n] (PEM%GlobalElemID), ParticleInside[n] and
 * IsNewPart[n] LOGICALs (particle_vars.f90:50-167); PartPosRef[n][3] or NULL; ids[n] or NULL. */
int piclas_gpu_upload_particles(int64_t n, const double *PartState, const int32_t *PartSpecies,
                                const int32_t *GlobalElemID, const int32_t *ParticleInside,
                                const int32_t *IsNewPart, const double *PartPosRef, const int64_t *ids,
                                int32_t append);

/* replaces the initial emission of one Part-Species$-Init$ whose SpaceIC is a lattice with a displacement in x and whose
 * velocityDistribution is constant (InitialParticleInserting -> SetParticlePosition -> SetParticlePositionSinDeviation /
 * SetParticlePositionCosDistribution, particle_emission_tools.f90:1235-1371; SinglePointToElement(doHALO=F),
 * particle_position_and_velocity.f90:434, particle_localization.f90:81-190; SetParticleVelocity 'constant').
 * SpaceIC: PGPU_EMIT_SIN_DEVIATION / PGPU_EMIT_COS_DISTRIBUTION; maxParticleNumber[3] = maxParticleNumberX/Y/Z;
 * velocity[3] = VeloIC * VeloVecIC.  Every rank calls it with the same arguments and keeps the positions that lie in its own
 * elements; IsNewPart is set; particle ids = position in the reference's loop nest (0-based).  Needs the FIBGM tables of the
 * mesh.  nEmitted: particles this rank accepted. */
#define PGPU_EMIT_SIN_DEVIATION 1
#define PGPU_EMIT_COS_DISTRIBUTION 2
int piclas_gpu_emit_lattice(int32_t SpaceIC, int32_t iSpec, const int32_t *maxParticleNumber, double Amplitude,
                            double WaveNumber, const double *velocity, int32_t append, int64_t *nEmitted);

/* replaces CALL Deposition() (pic_depo.f90:944-1018).
 * PartSource: LOCAL [nElems][N+1][N+1][N+1][4] == PS_N(iElem)%PartSource(1:4,i,j,k) packed by element;
 * NodeSource: [nUniqueGlobalNodes][4] (cell_volweight_mean only) or NULL.  Either may be NULL. */
int piclas_gpu_deposit(double *PartSource, double *NodeSource);

/* The HDG source term of the Poisson solve reads only the charge density, PS_N(iElem)%PartSource(4,i,j,k)
 * (equations/poisson/equation.f90:1043, CalcSourceHDG).  After piclas_gpu_deposit(NULL, NULL) (or deposit_finish) this
 * copies that one component, LOCAL [nElems][N+1][N+1][N+1], a quarter of the PartSource transfer. */
int piclas_gpu_get_charge(double *ChargeDensity);

/* PartSource of the last deposition to the host without stopping the step: the device -> host copy runs on its own stream, behind
 * the deposition and beside the calls that follow (the host -> device copy of piclas_gpu_set_field uses the other direction of the
 * link).  The HDG source term needs PS_N%PartSource(4,:) only (piclas_gpu_get_charge above, equations/poisson/equation.f90:1043);
 * the current density PartSource(1:3,:) is read by output and analysis (timedisc.f90:399-410), i.e. after the step.  PartSource:
 * LOCAL [nElems][N+1][N+1][N+1][4], page-locked host memory for a truly asynchronous copy.  The next piclas_gpu_deposit waits on
 * the device for the copy; the host reads the array after piclas_gpu_partsource_wait. */
int piclas_gpu_get_partsource_async(double *PartSource);
int piclas_gpu_partsource_wait(void);

/* replaces CalcKineticEnergy / CalcNumPartsOfSpec of the particle analysis (particle_analyze_tools.f90:709-842) so that
 * PartAnalyze.csv needs no particle download: Ekin[nSpecies] in Joule (0.5 m v^2 below RelativisticLimit = (1e6/299792458)^2 c^2,
 * (gamma-1) m c^2 above; times MacroParticleFactor), nPart[nSpecies] simulation particles.  Either may be NULL. */
int piclas_gpu_kinetic_energy(double *Ekin, int64_t *nPart);

/* after CALL HDG(time,iter): E == U_N(iElem)%E(1:3,i,j,k) packed, LOCAL [nElems][N+1][N+1][N+1][3] */
int piclas_gpu_set_field(const double *E);

/* replaces timedisc_TimeStepPoissonByBorisLeapfrog.f90:109-215: LastPartPos/LastGlobalElemID copy,
 * InterpolateFieldToParticle, push, PerformTracking, (MPI exchange, see below), UpdateNextFreePosition.
 * nLost: particles removed by tracking (NbrOfLostParticles). */
int piclas_gpu_push_track(double dt, int64_t iter, int32_t *nLost);

/* device SoA -> host AoS, compacted 1..n (PDM%ParticleVecLength == n, all ParticleInside);
 * call before PerformAnalyze / WriteStateToHDF5 / load balance.  Any output pointer may be NULL. */
int64_t piclas_gpu_num_particles(void);
int piclas_gpu_download_particles(int64_t nmax, double *PartState, int32_t *PartSpecies,
                                  int32_t *GlobalElemID, double *PartPosRef, int64_t *ids, int64_t *n_out);

/* ---- particle exchange between ranks (replaces particle_mpi.f90:202-1024, message layout :158-183) ------
 * nRanks > 1: push_track leaves the step open.  exchange_info extracts the emigrants of this rank, grouped by
 * destination rank, into a device buffer of PartCommSize doubles per particle; the transport (NCCL/MPI/P2P) is
 * the caller's; exchange_finish appends the immigrants and runs the one sort of the step
 * (UpdateNextFreePosition).  Both calls are mandatory after every push_track on every rank, also when nothing
 * migrates; deposit / download / the next push_track fail while the step is open. */
int piclas_gpu_exchange_info(int32_t *partCommSize, int64_t *nSendPerRank /*[nRanks]*/, void **devSendBuf);
int piclas_gpu_exchange_recv_buffer(int64_t nRecvTotal, void **devRecvBuf);
/* Optional, for transports that work on device memory throughout: after exchange_info, devSendCounts is int64[nRanks] ON THE
 * DEVICE (the counts of nSendPerRank), so that the count exchange (IRecvNbOfParticles / SendNbOfParticles) needs no host->device
 * copy; sendCap / recvCapDoubles are the current capacities of the send / receive buffers in doubles (a caller may wrap the
 * buffers once and slice). */
int piclas_gpu_exchange_device_info(void **devSendCounts, int64_t *sendCapDoubles, int64_t *recvCapDoubles);
int piclas_gpu_exchange_finish(int64_t nRecvTotal);

/* ---- cell_volweight_mean node halo (replaces pic_depo_method.f90:565-673) ------------------------------
 * deposit() leaves the rank-local NodeSource on the device; the caller sums it over ranks (an all-reduce
 * is deterministic for a fixed rank count but not rank-ordered; piclas_gpu_node_halo_info below is) and
 * hands the result back before PartSource is formed.  Single-rank runs never call these. */
int piclas_gpu_nodesource_device(void **devNodeSource /* double[nUniqueGlobalNodes][4] */);
int piclas_gpu_deposit_finish(double *PartSource, double *NodeSource);
/* The same halo on the compact list of the nodes that several ranks contribute to (elements of two ranks meet there, directly
 * or through a periodic partner; the reference exchanges exactly these, pic_depo.f90:298-571): after deposit(), devSend holds this
 * rank's nDoubles = 4 * nSharedNodes sums; the caller all-gathers them into devRecvAll ([nRanks][nDoubles], rank order) and calls
 * piclas_gpu_deposit_finish, which adds them up in rank order (deterministic).  Alternative to the full-array sum above. */
int piclas_gpu_node_halo_info(int64_t *nDoubles, void **devSend, void **devRecvAll);

/* Optional: run the library's kernels and copies on the caller's CUDA stream (cudaStream_t) instead of its own, so that the
 * caller's collectives (NCCL, CUDA-aware MPI) are ordered with them without host synchronisation. */
int piclas_gpu_set_stream(void *cudaStream);

/* ---- shape-function DOF halo (replaces pic_depo_method.f90:940-996, ShapeMapping Send/RecvBuffer) ---------------
 * Multi-rank runs: deposit() also forms the contributions of local particles to elements of other ranks.
 * nSend/nRecvElemsPerRank[nRanks] are element counts (fixed by the geometry at init), each element carries
 * doublesPerElem = 4*(N+1)^3 doubles; exchange devSend -> devRecv (rank-ordered blocks), then call
 * piclas_gpu_deposit_finish(PartSource, NULL), which adds the received blocks rank after rank. */
int piclas_gpu_sf_halo_info(int64_t *nSendElemsPerRank, int64_t *nRecvElemsPerRank, int32_t *doublesPerElem,
                            void **devSend, void **devRecv);

/* timing of the last call's kernels (ms, CUDA events on the launch stream) and launch count */
int piclas_gpu_last_timing(double *ms_kernels, int32_t *nLaunches);
/* CUDA-event durations (ms) of the phases of the most recent deposit / push_track call:
 * [0] deposition particle kernel, [1] node + DOF kernels, [2] interpolate+push+track kernel, [3] sort + permute.
 * The replaced code feeds the same sections into LB_DEPO_*, LB_INTERPOLATION+LB_PUSH+LB_TRACK, LB_UNFP
 * (loadbalance/loadbalance_timers.f90:70-266). */
int piclas_gpu_phase_timing(double *ms4);

#ifdef __cplusplus
}
#endif
#endif
