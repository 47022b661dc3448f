!==================================================================================================================================
! MOD_Particle_GPU — ISO_C_BINDING glue between PICLas and libpiclas_gpu.so (include/piclas_gpu.h).
!
! Not compiled in this repository (no Fortran compiler in the build image); this is the binding a PICLas maintainer adds to
! src/particles/ and the #if USE_GPU variant of the particle half of TimeStepPoissonByBorisLeapfrog
! (src/timedisc/timedisc_TimeStepPoissonByBorisLeapfrog.f90:93-279).  Style follows the only in-tree BIND(C) precedent,
! src/output/output.f90:33-48.  Every function returns 0 on success; on error the glue calls Abort(__STAMP__,...).
!==================================================================================================================================
MODULE MOD_Particle_GPU
USE ISO_C_BINDING
IMPLICIT NONE
PRIVATE

TYPE, BIND(C) :: pgpu_mesh_t            ! field order == include/piclas_gpu.h
  INTEGER(C_INT32_T) :: nGlobalElems, nSides, nNonUniqueNodes, nUniqueGlobalNodes
  INTEGER(C_INT32_T) :: NGeo, N, offsetElem, nElems, elemInfoSize, sideInfoSize
  TYPE(C_PTR) :: ElemInfo, SideInfo, NodeCoords, NodeInfo, ElemNodeID, ElemSideNodeID, ConcaveElemSide
  TYPE(C_PTR) :: XCL_NGeo, dXCL_NGeo, XiCL_NGeo, wBaryCL_NGeo, ElemBaryNGeo, ElemRadius2NGeo, XiEtaZetaBasis, slenXiEtaZetaBasis
  TYPE(C_PTR) :: xGP, wGP, wBary, Elem_xGP, ElemsJ
  INTEGER(C_INT32_T) :: nBCs
  TYPE(C_PTR) :: bc_kind, bc_alpha
  INTEGER(C_INT32_T) :: nPeriodicVectors
  TYPE(C_PTR) :: PeriodicVectors, Periodic_nNodes, Periodic_offsetNode, Periodic_Nodes
  INTEGER(C_INT32_T) :: nPeriodicNodesTotal
  TYPE(C_PTR) :: NodeVolume
  REAL(C_DOUBLE) :: FIBGMdeltas(3), xyzminglob(3), xyzmaxglob(3)
  INTEGER(C_INT32_T) :: FIBGMmin(3), FIBGMmax(3)
  TYPE(C_PTR) :: FIBGM_nElems, FIBGM_offsetElem, FIBGM_Element
  INTEGER(C_INT32_T) :: nFIBGMElemsTotal
  TYPE(C_PTR) :: ElemEpsOneCell, ElemToBCSides, SideBCMetrics
  INTEGER(C_INT32_T) :: nBCSidesTotal
  TYPE(C_PTR) :: SideType, SideNormVec, SideDistance, BaseVectors0, BaseVectors1, BaseVectors2, BaseVectorsScale
  TYPE(C_PTR) :: SFElemr2, ElemRadiusNGeo, ElemToBGM, BaseVectors3
END TYPE

TYPE, BIND(C) :: pgpu_params_t
  INTEGER(C_INT32_T) :: TrackingMethod, RefMappingGuess
  REAL(C_DOUBLE)     :: RefMappingEps
  INTEGER(C_INT32_T) :: CartesianPeriodic, TimeDiscMethod, DoInterpolation, DoDeposition, DepositionType
  REAL(C_DOUBLE)     :: externalField(6), c2_inv
  INTEGER(C_INT32_T) :: nSpecies
  TYPE(C_PTR)        :: ChargeIC, MassIC, MacroParticleFactor
  REAL(C_DOUBLE)     :: r_sf
  INTEGER(C_INT32_T) :: alpha_sf, dim_sf, dim_sf_dir, sfDepo3D
  REAL(C_DOUBLE)     :: w_sf, dimFactorSF
  INTEGER(C_INT32_T) :: device, myRank, nRanks
  INTEGER(C_INT64_T) :: maxParticleNumber
  INTEGER(C_INT32_T) :: carryParticleIDs, arithmetic
  INTEGER(C_INT32_T) :: PartLorentzType, NoDirichletDeposition, DoDielectricSurfaceCharge   ! must be 0: not implemented
END TYPE

INTERFACE
  FUNCTION piclas_gpu_init(mesh,params) BIND(C,NAME='piclas_gpu_init')
    IMPORT :: C_INT, pgpu_mesh_t, pgpu_params_t
    TYPE(pgpu_mesh_t),INTENT(IN)   :: mesh
    TYPE(pgpu_params_t),INTENT(IN) :: params
    INTEGER(C_INT)                 :: piclas_gpu_init
  END FUNCTION
  FUNCTION piclas_gpu_finalize() BIND(C,NAME='piclas_gpu_finalize')
    IMPORT :: C_INT
    INTEGER(C_INT) :: piclas_gpu_finalize
  END FUNCTION
  FUNCTION piclas_gpu_last_error() BIND(C,NAME='piclas_gpu_last_error')
    IMPORT :: C_PTR
    TYPE(C_PTR) :: piclas_gpu_last_error
  END FUNCTION
  FUNCTION piclas_gpu_upload_particles(n,PartState,PartSpecies,GlobalElemID,ParticleInside,IsNewPart,PartPosRef,ids,append) &
      BIND(C,NAME='piclas_gpu_upload_particles')
    IMPORT :: C_INT, C_INT32_T, C_INT64_T, C_DOUBLE, C_PTR
    INTEGER(C_INT64_T),VALUE :: n
    REAL(C_DOUBLE),INTENT(IN)     :: PartState(6,*)
    INTEGER(C_INT32_T),INTENT(IN) :: PartSpecies(*),GlobalElemID(*),ParticleInside(*),IsNewPart(*)
    TYPE(C_PTR),VALUE             :: PartPosRef, ids           ! C_NULL_PTR unless RefMapping / id tracking
    INTEGER(C_INT32_T),VALUE      :: append
    INTEGER(C_INT)                :: piclas_gpu_upload_particles
  END FUNCTION
  FUNCTION piclas_gpu_deposit(PartSource,NodeSource) BIND(C,NAME='piclas_gpu_deposit')
    IMPORT :: C_INT, C_PTR
    TYPE(C_PTR),VALUE :: PartSource, NodeSource                  ! packed [4,nDOF_local] / [4,nUniqueGlobalNodes] or C_NULL_PTR
    INTEGER(C_INT)    :: piclas_gpu_deposit
  END FUNCTION
  FUNCTION piclas_gpu_get_charge(ChargeDensity) BIND(C,NAME='piclas_gpu_get_charge')
    IMPORT :: C_INT, C_DOUBLE
    REAL(C_DOUBLE),INTENT(OUT) :: ChargeDensity(*)               ! packed [nDOF_local] == PS_N(iElem)%PartSource(4,i,j,k)
    INTEGER(C_INT)             :: piclas_gpu_get_charge
  END FUNCTION
  FUNCTION piclas_gpu_get_partsource_async(PartSource) BIND(C,NAME='piclas_gpu_get_partsource_async')
    IMPORT :: C_INT, C_DOUBLE
    REAL(C_DOUBLE),INTENT(OUT) :: PartSource(4,*)                ! packed [4,nDOF_local]; valid after piclas_gpu_partsource_wait
    INTEGER(C_INT)             :: piclas_gpu_get_partsource_async
  END FUNCTION
  FUNCTION piclas_gpu_partsource_wait() BIND(C,NAME='piclas_gpu_partsource_wait')
    IMPORT :: C_INT
    INTEGER(C_INT)             :: piclas_gpu_partsource_wait
  END FUNCTION
  FUNCTION piclas_gpu_kinetic_energy(Ekin,nPart) BIND(C,NAME='piclas_gpu_kinetic_energy')
    IMPORT :: C_INT, C_DOUBLE, C_INT64_T
    REAL(C_DOUBLE),INTENT(OUT)     :: Ekin(*)                    ! [nSpecies] -> CalcKineticEnergy
    INTEGER(C_INT64_T),INTENT(OUT) :: nPart(*)                   ! [nSpecies] -> CalcNumPartsOfSpec
    INTEGER(C_INT)                 :: piclas_gpu_kinetic_energy
  END FUNCTION
  FUNCTION piclas_gpu_emit_lattice(SpaceIC,iSpec,maxParticleNumber,Amplitude,WaveNumber,velocity,append,nEmitted) &
      BIND(C,NAME='piclas_gpu_emit_lattice')
    IMPORT :: C_INT, C_INT32_T, C_INT64_T, C_DOUBLE
    INTEGER(C_INT32_T),VALUE       :: SpaceIC                    ! 1 sin_deviation, 2 cos_distribution
    INTEGER(C_INT32_T),VALUE       :: iSpec
    INTEGER(C_INT32_T),INTENT(IN)  :: maxParticleNumber(3)       ! Species(iSpec)%Init(iInit)%maxParticleNumberX/Y/Z
    REAL(C_DOUBLE),VALUE           :: Amplitude, WaveNumber
    REAL(C_DOUBLE),INTENT(IN)      :: velocity(3)                ! VeloIC*VeloVecIC
    INTEGER(C_INT32_T),VALUE       :: append
    INTEGER(C_INT64_T),INTENT(OUT) :: nEmitted
    INTEGER(C_INT)                 :: piclas_gpu_emit_lattice
  END FUNCTION
  FUNCTION piclas_gpu_set_field(E) BIND(C,NAME='piclas_gpu_set_field')
    IMPORT :: C_INT, C_DOUBLE
    REAL(C_DOUBLE),INTENT(IN) :: E(3,*)                          ! packed [3,nDOF_local]
    INTEGER(C_INT)            :: piclas_gpu_set_field
  END FUNCTION
  FUNCTION piclas_gpu_push_track(dt,iter,nLost) BIND(C,NAME='piclas_gpu_push_track')
    IMPORT :: C_INT, C_INT32_T, C_INT64_T, C_DOUBLE
    REAL(C_DOUBLE),VALUE     :: dt
    INTEGER(C_INT64_T),VALUE :: iter
    INTEGER(C_INT32_T),INTENT(OUT) :: nLost
    INTEGER(C_INT)           :: piclas_gpu_push_track
  END FUNCTION
  FUNCTION piclas_gpu_num_particles() BIND(C,NAME='piclas_gpu_num_particles')
    IMPORT :: C_INT64_T
    INTEGER(C_INT64_T) :: piclas_gpu_num_particles
  END FUNCTION
  FUNCTION piclas_gpu_download_particles(nmax,PartState,PartSpecies,GlobalElemID,PartPosRef,ids,n_out) &
      BIND(C,NAME='piclas_gpu_download_particles')
    IMPORT :: C_INT, C_INT32_T, C_INT64_T, C_DOUBLE, C_PTR
    INTEGER(C_INT64_T),VALUE :: nmax
    REAL(C_DOUBLE),INTENT(OUT)     :: PartState(6,*)
    INTEGER(C_INT32_T),INTENT(OUT) :: PartSpecies(*),GlobalElemID(*)
    TYPE(C_PTR),VALUE              :: PartPosRef, ids
    INTEGER(C_INT64_T),INTENT(OUT) :: n_out
    INTEGER(C_INT)                 :: piclas_gpu_download_particles
  END FUNCTION
  ! ---- multi-rank: particle migration (particle_mpi.f90:202-1024), device buffers for CUDA-aware MPI / NCCL
  FUNCTION piclas_gpu_exchange_info(partCommSize,nSendPerRank,devSendBuf) BIND(C,NAME='piclas_gpu_exchange_info')
    IMPORT :: C_INT, C_INT32_T, C_INT64_T, C_PTR
    INTEGER(C_INT32_T),INTENT(OUT) :: partCommSize
    INTEGER(C_INT64_T),INTENT(OUT) :: nSendPerRank(*)            ! [nProcessors]
    TYPE(C_PTR),INTENT(OUT)        :: devSendBuf
    INTEGER(C_INT)                 :: piclas_gpu_exchange_info
  END FUNCTION
  FUNCTION piclas_gpu_exchange_recv_buffer(nRecvTotal,devRecvBuf) BIND(C,NAME='piclas_gpu_exchange_recv_buffer')
    IMPORT :: C_INT, C_INT64_T, C_PTR
    INTEGER(C_INT64_T),VALUE :: nRecvTotal
    TYPE(C_PTR),INTENT(OUT)  :: devRecvBuf
    INTEGER(C_INT)           :: piclas_gpu_exchange_recv_buffer
  END FUNCTION
  FUNCTION piclas_gpu_exchange_device_info(devSendCounts,sendCapDoubles,recvCapDoubles) BIND(C,NAME='piclas_gpu_exchange_device_info')
    IMPORT :: C_INT, C_INT64_T, C_PTR
    TYPE(C_PTR),INTENT(OUT)        :: devSendCounts              ! INTEGER(8) nSendPerRank(1:nProcessors) on the device
    INTEGER(C_INT64_T),INTENT(OUT) :: sendCapDoubles,recvCapDoubles
    INTEGER(C_INT)                 :: piclas_gpu_exchange_device_info
  END FUNCTION
  FUNCTION piclas_gpu_exchange_finish(nRecvTotal) BIND(C,NAME='piclas_gpu_exchange_finish')
    IMPORT :: C_INT, C_INT64_T
    INTEGER(C_INT64_T),VALUE :: nRecvTotal
    INTEGER(C_INT)           :: piclas_gpu_exchange_finish
  END FUNCTION
  ! ---- multi-rank deposition: node halo of cell_volweight_mean (pic_depo_method.f90:565-673), DOF halo of the shape functions (:940-996)
  FUNCTION piclas_gpu_nodesource_device(devNodeSource) BIND(C,NAME='piclas_gpu_nodesource_device')
    IMPORT :: C_INT, C_PTR
    TYPE(C_PTR),INTENT(OUT) :: devNodeSource                     ! [4,nUniqueGlobalNodes] doubles on the device
    INTEGER(C_INT)          :: piclas_gpu_nodesource_device
  END FUNCTION
  FUNCTION piclas_gpu_node_halo_info(nDoubles,devSend,devRecvAll) BIND(C,NAME='piclas_gpu_node_halo_info')
    IMPORT :: C_INT, C_INT64_T, C_PTR
    INTEGER(C_INT64_T),INTENT(OUT) :: nDoubles                   ! 4 * number of nodes shared between ranks
    TYPE(C_PTR),INTENT(OUT)        :: devSend,devRecvAll         ! device buffers [nDoubles], [nProcessors*nDoubles] for MPI_ALLGATHER
    INTEGER(C_INT)                 :: piclas_gpu_node_halo_info
  END FUNCTION
  FUNCTION piclas_gpu_set_stream(cudaStream) BIND(C,NAME='piclas_gpu_set_stream')
    IMPORT :: C_INT, C_PTR
    TYPE(C_PTR),VALUE :: cudaStream                              ! cudaStream_t of the host's communication library
    INTEGER(C_INT)    :: piclas_gpu_set_stream
  END FUNCTION
  FUNCTION piclas_gpu_sf_halo_info(nSendElemsPerRank,nRecvElemsPerRank,doublesPerElem,devSend,devRecv) &
      BIND(C,NAME='piclas_gpu_sf_halo_info')
    IMPORT :: C_INT, C_INT32_T, C_INT64_T, C_PTR
    INTEGER(C_INT64_T),INTENT(OUT) :: nSendElemsPerRank(*),nRecvElemsPerRank(*)
    INTEGER(C_INT32_T),INTENT(OUT) :: doublesPerElem
    TYPE(C_PTR),INTENT(OUT)        :: devSend,devRecv
    INTEGER(C_INT)                 :: piclas_gpu_sf_halo_info
  END FUNCTION
  FUNCTION piclas_gpu_deposit_finish(PartSource,NodeSource) BIND(C,NAME='piclas_gpu_deposit_finish')
    IMPORT :: C_INT, C_PTR
    TYPE(C_PTR),VALUE :: PartSource, NodeSource
    INTEGER(C_INT)    :: piclas_gpu_deposit_finish
  END FUNCTION
  FUNCTION piclas_gpu_phase_timing(ms4) BIND(C,NAME='piclas_gpu_phase_timing')
    IMPORT :: C_INT, C_DOUBLE
    REAL(C_DOUBLE),INTENT(OUT) :: ms4(4)                         ! -> LBSplitTime(LB_DEPO / LB_INTERPOLATION+LB_PUSH+LB_TRACK / LB_UNFP)
    INTEGER(C_INT)             :: piclas_gpu_phase_timing
  END FUNCTION
  FUNCTION piclas_gpu_last_timing(ms_kernels,nLaunches) BIND(C,NAME='piclas_gpu_last_timing')
    IMPORT :: C_INT, C_INT32_T, C_DOUBLE
    REAL(C_DOUBLE),INTENT(OUT)     :: ms_kernels                 ! CUDA-event time of the last call's kernels
    INTEGER(C_INT32_T),INTENT(OUT) :: nLaunches                  ! kernels launched by the last call
    INTEGER(C_INT)                 :: piclas_gpu_last_timing
  END FUNCTION
END INTERFACE

PUBLIC :: pgpu_mesh_t, pgpu_params_t
PUBLIC :: piclas_gpu_init, piclas_gpu_finalize, piclas_gpu_upload_particles, piclas_gpu_deposit, piclas_gpu_set_field
PUBLIC :: piclas_gpu_push_track, piclas_gpu_num_particles, piclas_gpu_download_particles, piclas_gpu_get_charge, piclas_gpu_kinetic_energy
PUBLIC :: piclas_gpu_exchange_info, piclas_gpu_exchange_recv_buffer, piclas_gpu_exchange_finish
PUBLIC :: piclas_gpu_nodesource_device, piclas_gpu_sf_halo_info, piclas_gpu_deposit_finish, piclas_gpu_phase_timing
PUBLIC :: piclas_gpu_last_timing, piclas_gpu_node_halo_info, piclas_gpu_set_stream, piclas_gpu_exchange_device_info
PUBLIC :: piclas_gpu_emit_lattice, piclas_gpu_get_partsource_async, piclas_gpu_partsource_wait
PUBLIC :: ParticleStepGPU, GPUAbortOnError

CONTAINS

SUBROUTINE GPUAbortOnError(rc,stamp)
! mirrors every CALL abort(__STAMP__,...) of the replaced code (globals/globals.f90:322-397)
USE MOD_Globals ,ONLY: abort
INTEGER(C_INT),INTENT(IN)   :: rc
CHARACTER(LEN=*),INTENT(IN) :: stamp
CHARACTER(KIND=C_CHAR),POINTER :: msg(:)
CHARACTER(LEN=1024) :: text
INTEGER :: i
IF (rc.EQ.0) RETURN
CALL C_F_POINTER(piclas_gpu_last_error(),msg,(/1024/))
text=''
DO i=1,1024
  IF (msg(i).EQ.C_NULL_CHAR) EXIT
  text(i:i)=msg(i)
END DO
CALL abort(__STAMP__,TRIM(stamp)//': '//TRIM(text))
END SUBROUTINE GPUAbortOnError

!==================================================================================================================================
! Particle half of TimeStepPoissonByBorisLeapfrog with the device layer (replaces :93, :109-215, :270 of the original).
! PS_N / U_N are packed through N_DG_Mapping(1,elem) offsets exactly like Elem_xGP_Shared (particle_mesh_tools.f90:1586-1593).
!==================================================================================================================================
SUBROUTINE ParticleStepGPU(PartSourcePacked,EPacked,dt,iter)
USE MOD_HDG ,ONLY: HDG
REAL(C_DOUBLE),INTENT(INOUT),TARGET :: PartSourcePacked(:,:) ! (4,nDOF_local)
REAL(C_DOUBLE),INTENT(INOUT)        :: EPacked(:,:)          ! (3,nDOF_local)
REAL,INTENT(IN)                     :: dt
INTEGER(KIND=8),INTENT(IN)          :: iter
INTEGER(C_INT32_T) :: nLost
REAL               :: time
time=0.
! Deposition()  -> PartSource on the host for the HDG right-hand side (equations/poisson/equation.f90:1043)
CALL GPUAbortOnError(piclas_gpu_deposit(C_LOC(PartSourcePacked),C_NULL_PTR),'piclas_gpu_deposit')
! ... unpack PartSourcePacked into PS_N(iElem)%PartSource, CALL HDG(time,iter), pack U_N(iElem)%E into EPacked ...
CALL GPUAbortOnError(piclas_gpu_set_field(EPacked),'piclas_gpu_set_field')
! LastPartPos/LastGlobalElemID copy, InterpolateFieldToParticle, push, PerformTracking, UpdateNextFreePosition
CALL GPUAbortOnError(piclas_gpu_push_track(REAL(dt,C_DOUBLE),INT(iter,C_INT64_T),nLost),'piclas_gpu_push_track')
! NbrOfLostParticles = NbrOfLostParticles + nLost   (particle_tracking_vars)
END SUBROUTINE ParticleStepGPU

END MODULE MOD_Particle_GPU
