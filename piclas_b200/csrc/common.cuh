// common.cuh — device-side data layout of the B200 particle step (see DESIGN.md "Data layout in HBM").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/piclas_gpu.h"

#define PGPU_MAX_N 7          // highest supported solution degree (N+1 <= 8)
#define PGPU_NCORNER 8        // NGeo == 1

// ---- per-element records, built once at init from the host tables ---------------------------------------------
// Tria record: everything ParticleInsideQuad3D / ThroughSideCheck3DFast / IntersectionWithWall need for one
// element (reference tables ElemInfo/SideInfo/ElemSideNodeID/NodeCoords/ConcaveElemSide),
// flattened so that a CTA can stage it with one coalesced copy.  352 B; a corner is one aligned 32-byte unit (256-bit loads).
struct __align__(32) TriaElem {
  double corner[8][4];     // the element's 8 non-unique nodes in storage (tensor) order: NodeCoords(:,first+1..first+8), [3] = pad
  int32_t nbElem[6];       // SIDE_NBELEMID of local side 1..6 (global id, 0 = none)
  int32_t sideID[6];       // global SideInfo index (1-based) of local side 1..6
  uint8_t sideNode[6][4];  // ElemSideNodeID(1:4,side) - ELEM_FIRSTNODEIND  (0..7)
  uint8_t bcid[6];         // SIDE_BCID (0 = inner side); nBCs <= 255
  uint8_t concave;         // bit s-1 = ConcaveElemSide(s)
  uint8_t pad[1];
};

// Geometry record for the Newton reference mapping at NGeo == 1: 123 doubles.
struct __align__(16) GeoElem {
  double XCL[8][3];        // XCL_NGeo(1:3,i,j,k), node index i+2j+4k
  double dXCL[8][3][3];    // [node][nn][dd] == dXCL_NGeo(dd,nn,i,j,k)
  double bary[3];          // ElemBaryNGeo
  double xez[6][3];        // XiEtaZetaBasis(1:3,1:6)
  double slen[6];          // slenXiEtaZetaBasis(1:6)
  double pad;              // keep sizeof a multiple of 16
};

// Fast-arithmetic records (params.arithmetic == 1), derived at init from the tables above.
// PlaneElem: inward unit normal and offset of the 12 side triangles: the determinant of ParticleInsideQuad3D for triangle
// t equals (n_t . x - d_t) * |N_t|, so its sign is known without evaluating it whenever |n_t . x - d_t| > tol.
struct __align__(32) PlaneElem {
  double pl[12][4];  // index 2*s + (tri-1): (n_x, n_y, n_z, d), one aligned 32-byte unit per plane
  double tol;        // 1e-8 * element diameter: far above the rounding error of either formula
  uint32_t concave2; // bit 2*s set when ConcaveElemSide(s+1)
  uint32_t planar;   // 1: convex element whose six sides are planar (both triangles in one plane): exit-side shortcut allowed
  uint32_t pad[4];
  double dg[6][4];   // per side: plane through the triangle diagonal (node 1 -> node 3) normal to the side, positive towards
                     // node 2, i.e. towards triangle 1; decides which triangle of a planar side a crossing point lies in
};
// AffElem: elements whose trilinear map is affine (parallelepipeds): xi = A (x - x0) - 1 solves the Newton problem exactly
struct __align__(16) AffElem {
  double x0[3];
  double A[3][3];
  double affine;     // 1.0 if the shortcut applies
  double pad[3];
};

// small read-only tables in constant memory
struct ConstTables {
  double xGP[PGPU_MAX_N + 1], wGP[PGPU_MAX_N + 1], wBary[PGPU_MAX_N + 1];
  double cvwFac[PGPU_MAX_N + 1];   // CellVolWeight%Fac = (xGP+1)/2
  double XiCL[2], wBaryCL[2];
  double externalField[6];
  double c2_inv;
  double RefMappingEps;
  int32_t RefMappingGuess;
  int32_t TrackingMethod, TimeDiscMethod, DoInterpolation, DepositionType;
  int32_t nSpecies;
  double ChargeIC[32], MassIC[32], MPF[32];
  int32_t nBCs;
  int32_t bc_kind[256], bc_alpha[256];
  int32_t nPeriodicVectors;
  double PeriodicVectors[8][3];
  int32_t nGlobalElems, nElems, offsetElem, N, nRanks, myRank;
  int32_t arithmetic;
  // shape function (pic_depo.f90:828-910, pic_depo_shapefunction_tools.f90:1131-1294)
  int32_t dim_sf, dim_sf_dir, dim_sf_dir1, dim_sf_dir2, dim_periodic_vec1, dim_periodic_vec2, nSFCases, alpha_sf, sfDepo3D;
  int32_t sfCase[27][3];
  double r_sf, r2_sf, r2_sf_inv, w_sf, dimFactorSF;
  double FIBGMdeltas[3], xyzminglob[3];
  // restructured arithmetic: nodal (Gauss) values -> monomial coefficients, a_p = sum_i n2m[p][i] u_i (inverse Vandermonde)
  double n2m[PGPU_MAX_N + 1][PGPU_MAX_N + 1];
  int32_t FIBGMmin[3], FIBGMmax[3];
};

// particle SoA (one of two buffers)
struct PartBuf {
  double *f;          // one block [6][stride]: x0,x1,x2,v0,v1,v2 (x[d] and v[d] below point into it)
  double *xif;        // one block [3][stride] behind xi[d]
  int64_t stride;
  double *x[3];
  double *v[3];
  double *xi[3];      // cached reference position of the current (x, elem); valid iff Ctx::xiValid
  int32_t *elem;      // PEM%GlobalElemID (1-based global), 0 = removed
  uint8_t *meta;      // bits 0-5 species-1, bit 6 = xi mapping failed (SucRefPos = F), bit 7 = PDM%IsNewPart
  int64_t *id;        // optional
};

#define META_SPEC_MASK 0x3F
#define META_XIFAIL 0x40
#define META_ISNEW 0x80
#define META_UNPUSHED 0x40   /* far records of the binned layout only (same bit as META_XIFAIL, which the bins do not use) */

#define KEY_DEAD 0xFFFFFFFFu
