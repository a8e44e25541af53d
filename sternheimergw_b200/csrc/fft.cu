// fft.cu -- batched, sphere-pruned 3-D complex-FP64 FFT pipeline for the local-potential part of H.psi,
// dV.psi (dvqpsi_us.f90:99-130) and the Delta-rho accumulation ([QE] vloc_psi_k / incdrhoscf semantics).
//
// Decomposition (per vector):   z-pass over sphere columns  ->  fused 2-D plane kernel  ->  z-pass
//   k_zpass_g2r : gather a chunk of (x,y) columns of the sphere into shared memory, inverse 1-D FFT along z,
//                 write T[vec][pz][col]                    (only the ~pi r^2 non-empty columns are touched)
//   k_plane     : one CTA per (vec, z-plane): scatter the columns into an nx*ny shared-memory plane, inverse
//                 FFT along y (only x-columns that hold data) and x, multiply by v(r) (or a complex field),
//                 forward FFT along x and y (only needed columns), gather the output sphere's columns
//   k_zpass_r2g : forward 1-D FFT along z per column, scale 1/nnr, fused H.psi epilogue
//                 (kinetic + sigma*psi + non-local part already in `out`)
// Real space is held in the radix-permuted order of fft_core.h; nothing leaves the chip between the inverse
// and the forward 2-D transform, so per vector the HBM traffic is ~4 x ncol*nz*16 B instead of the 12 full
// box sweeps of an unfused 3-D FFT (SURVEY.md section 8d counts 192*nnr + 32*N algorithmic bytes).
#include "internal.cuh"

#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

namespace sgw {

constexpr int ZCB_DEFAULT = 16; // columns per CTA in the z passes (template parameter ZCB of the kernels)
constexpr int ZTHREADS = 256;   // launch bound; the launch uses zpass_threads()
constexpr int ZMINB16 = 6;      // resident CTAs per SM the register allocation of the ZCB <= 16 z-pass kernels must allow

constexpr int PTHREADS = 256;

template <int ZCB>
__global__ void __launch_bounds__(ZCB >= 32 ? 256 : 160, ZCB >= 32 ? 4 : ZMINB16) k_zpass_g2r(GridDev g, SphereDev s, const cplx *__restrict__ in, long ld,
                                                         cplx *__restrict__ T, const int *__restrict__ active) {
  const int vec = blockIdx.y;
  if (active && !active[vec]) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int c0 = blockIdx.x * ZCB;
  const int nc = min(ZCB, s.ncol - c0);
  extern __shared__ cplx sm[];
  const int pitch = g.nz | 1;
  cplx *lines = sm;
  cplx *tw = sm + ZCB * pitch;
  for (int i = tid; i < ZCB * pitch; i += nt) lines[i] = cmake(0.0, 0.0);
  for (int i = tid; i < g.nz; i += nt) tw[i] = g.twz[i];
  __syncthreads();
  const int p0 = s.col_ptr[c0], p1 = s.col_ptr[c0 + nc];
  const cplx *src = in + (long)vec * ld;
  for (int p = p0 + tid; p < p1; p += nt) lines[(s.colof[p] - c0) * pitch + s.zof[p]] = src[p];
  __syncthreads();
  run_strided<+1>(g.rz1, lines, nc, nullptr, pitch, 1, g.rz2, tw, g.rz2 > 1, tid, nt);
  __syncthreads();
  if (g.rz2 > 1) {
    run_contig<+1>(g.rz2, lines, nc, nullptr, pitch, 1, g.rz1, tw, false, tid, nt);
    __syncthreads();
  }
  cplx *dst = T + (long)vec * g.nz * s.ncol + c0;
  for (int i = tid; i < ZCB * g.nz; i += nt) {
    const int c = i & (ZCB - 1), pz = i / ZCB;
    if (c < nc) dst[pz * s.ncol + c] = lines[c * pitch + pz];
  }
}

// RZ1 x RZ2 > 0: the radix plan of the z axis is a compile-time constant (the stages inline, no dispatch call)
template <int ZCB, int RZ1 = 0, int RZ2 = 0>
__global__ void __launch_bounds__(ZCB >= 32 ? 256 : 160, ZCB >= 32 ? 4 : ZMINB16) k_zpass_r2g(GridDev g, SphereDev s, const cplx *__restrict__ T,
                                                         cplx *__restrict__ out, long ld, ZEpilogue epi, double scale,
                                                         const int *__restrict__ active) {
  const int vec = blockIdx.y;
  if (active && !active[vec]) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int c0 = blockIdx.x * ZCB;
  const int nc = min(ZCB, s.ncol - c0);
  extern __shared__ cplx sm[];
  const int pitch = g.nz | 1;
  cplx *lines = sm;
  cplx *tw = sm + ZCB * pitch;
  for (int i = tid; i < g.nz; i += nt) tw[i] = g.twz[i];
  const cplx *src = T + (long)vec * g.nz * s.ncol + c0;
  for (int i = tid; i < ZCB * g.nz; i += nt) {
    const int c = i & (ZCB - 1), pz = i / ZCB;
    if (c < nc) lines[c * pitch + pz] = src[pz * s.ncol + c];
  }
  __syncthreads();
  if (RZ1 > 0) {
    constexpr int R1 = (RZ1 > 0) ? RZ1 : 1, R2 = (RZ2 > 0) ? RZ2 : 1;
    stage_contig<R2, -1>(lines, nc, nullptr, pitch, 1, R1, tw, true, tid, nt);
    __syncthreads();
    stage_strided<R1, -1>(lines, nc, nullptr, pitch, 1, R2, tw, false, tid, nt);
  } else {
    if (g.rz2 > 1) {
      run_contig<-1>(g.rz2, lines, nc, nullptr, pitch, 1, g.rz1, tw, true, tid, nt);
      __syncthreads();
    }
    run_strided<-1>(g.rz1, lines, nc, nullptr, pitch, 1, g.rz2, tw, false, tid, nt);
  }
  __syncthreads();
  const int p0 = s.col_ptr[c0], p1 = s.col_ptr[c0 + nc];
  cplx *dst = out + (long)vec * ld;
  cplx sg = cmake(0.0, 0.0);
  if (epi.mode == 1 && epi.sigma) sg = epi.sigma[(long)vec * epi.sigma_stride];
  for (int p = p0 + tid; p < p1; p += nt) {
    cplx val = cscale(scale, lines[(s.colof[p] - c0) * pitch + s.zof[p]]);
    if (epi.mode == 0) {
      dst[p] = val;
    } else if (epi.mode == 1) {
      const cplx ps = epi.psi[(long)vec * ld + p];
      cplx o = epi.keep_out ? dst[p] : cmake(0.0, 0.0);
      // (H + sigma) psi = g2kin psi + V_loc psi + [non-local, already in o] + sigma psi
      val = cadd(cscale(epi.g2kin[p], ps), val);
      val = cadd(val, o);
      dst[p] = cfma(sg, ps, val);
    } else {
      dst[p] = cadd(dst[p], val);
    }
  }
}

// GMEM: the plane does not fit in shared memory (boxes beyond ~118 x 118): it lives in a per-CTA slice of `scratch` in global
// memory (L2 for the sizes in question) and the same stages run on it; `vec0` is the first vector of the launch (the host
// walks over the batch in chunks that bound the scratch).  Slower per point, same arithmetic.
template <int MODE, int NT, bool ONE, bool GMEM = false>
__global__ void __launch_bounds__(NT, 2) k_plane(GridDev g, SphereDev sin, SphereDev sout, const cplx *__restrict__ Tin,
                                                     cplx *__restrict__ Tout, const double *__restrict__ vperm,
                                                     const cplx *__restrict__ field, int vec_per_field, cplx *R,
                                                     int in_mod, const int *__restrict__ active, int Prt,
                                                     cplx *scratch = nullptr, int vec0 = 0) {
  const int P = ONE ? 1 : Prt;        // ONE: single-plane instantiation (large boxes), plane loops fold away
  // One CTA = P consecutive z-planes of one vector, stacked in shared memory (plane p at offset p * ny * pitch, so the
  // rows of all planes form ONE set of np * ny lines with stride `pitch` and the potential rows v[pz0 * nxy + l * nx]
  // stay contiguous).  P = 1 for large boxes (72^2 fills the CTA); small boxes (15^2..24^2) pack up to 8 planes so that
  // the radix stages have enough butterflies for every thread.
  // in_mod > 0: the INPUT (Tin or R) of vector `vec` is input vector vec % in_mod (one set of bands shared by
  // every perturbation, dvqpsi_us.f90:99-130)
  const int vec = blockIdx.y + (GMEM ? vec0 : 0), pz0 = blockIdx.x * P;
  const int np = ONE ? 1 : min(P, g.nz - pz0);
  const int vin = in_mod > 0 ? vec % in_mod : vec;
  if (active && !active[vec]) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  extern __shared__ cplx sm[];
  const int pitch = g.pitchx, nx = g.nx, ny = g.ny;
  const int psz = ny * pitch;
  cplx *plane = GMEM ? scratch + ((long)blockIdx.y * gridDim.x + blockIdx.x) * psz : sm;
  cplx *twx = GMEM ? sm : sm + P * psz;
  cplx *twy = twx + nx;
  int *ids_in = (int *)(twy + ny), *ids_out = ids_in + P * nx;   // line starts of the x columns that hold data, all planes
  for (int i = tid; i < nx; i += nt) twx[i] = g.twx[i];
  for (int i = tid; i < ny; i += nt) twy[i] = g.twy[i];
  if (MODE != PLANE_FROM_R)
    for (int i = tid; i < np * sin.nxs; i += nt) ids_in[i] = ONE ? sin.xs[i] : (i / sin.nxs) * psz + sin.xs[i % sin.nxs];
  if (MODE != PLANE_TO_R)
    for (int i = tid; i < np * sout.nxs; i += nt) ids_out[i] = ONE ? sout.xs[i] : (i / sout.nxs) * psz + sout.xs[i % sout.nxs];
  const long nxy = (long)nx * ny;
  const int nrows = np * ny;

  if (MODE != PLANE_FROM_R) {
    for (int i = tid; i < np * psz; i += nt) plane[i] = cmake(0.0, 0.0);
    __syncthreads();
    const cplx *row = Tin + ((long)vin * g.nz + pz0) * sin.ncol;
    for (int i = tid; i < np * sin.ncol; i += nt) {
      const int p = ONE ? 0 : i / sin.ncol, c = i - p * sin.ncol;
      plane[p * psz + sin.col_off[c]] = row[i];
    }
    __syncthreads();
    // inverse along y for the x columns that hold data (x still in natural order)
    run_strided<+1>(g.ry1, plane, np * sin.nxs, ids_in, 1, pitch, g.ry2, twy, g.ry2 > 1, tid, nt);
    __syncthreads();
    if (g.ry2 > 1) {
      run_contig<+1>(g.ry2, plane, np * sin.nxs, ids_in, 1, pitch, g.ry1, twy, false, tid, nt);
      __syncthreads();
    }
    if (MODE == PLANE_VLOC || MODE == PLANE_FIELD) {
      // inverse along x, x v(r), forward along x: the last inverse stage, the product and the first forward stage
      // touch the same contiguous radix group of a row -> one register round trip (fft_core.h stage_mid)
      const double *v = MODE == PLANE_VLOC ? vperm + (long)pz0 * nxy : nullptr;
      const cplx *f = MODE == PLANE_FIELD ? field + ((long)(vec / vec_per_field) * g.nz + pz0) * nxy : nullptr;
      if (g.rx2 > 1) {
        run_strided<+1>(g.rx1, plane, nrows, nullptr, pitch, 1, g.rx2, twx, true, tid, nt);
        __syncthreads();
        run_mid<MODE == PLANE_FIELD>(g.rx2, plane, nrows, pitch, g.rx1, twx, true, v, f, nx, tid, nt);
        __syncthreads();
        run_strided<-1>(g.rx1, plane, nrows, nullptr, pitch, 1, g.rx2, twx, false, tid, nt);
      } else {
        run_mid<MODE == PLANE_FIELD>(g.rx1, plane, nrows, pitch, 1, twx, false, v, f, nx, tid, nt);
      }
      __syncthreads();
    } else {
      // inverse along x for all rows
      run_strided<+1>(g.rx1, plane, nrows, nullptr, pitch, 1, g.rx2, twx, g.rx2 > 1, tid, nt);
      __syncthreads();
      if (g.rx2 > 1) {
        run_contig<+1>(g.rx2, plane, nrows, nullptr, pitch, 1, g.rx1, twx, false, tid, nt);
        __syncthreads();
      }
    }
  } else {
    __syncthreads();
  }

  if (MODE == PLANE_VLOC || MODE == PLANE_FIELD) {
    // x transforms and the product are done (fused above)
  } else if (MODE == PLANE_TO_R) {
    cplx *r = R + ((long)vec * g.nz + pz0) * nxy;
    for (int i = tid; i < nx * nrows; i += nt) {
      const int ix = i % nx, iy = i / nx;
      r[i] = plane[iy * pitch + ix];
    }
    return;
  } else {  // PLANE_FROM_R (optionally times a complex field: dV_bare(r) psi(r))
    const cplx *r = R + ((long)vin * g.nz + pz0) * nxy;
    if (field) {
      const cplx *f = field + ((long)(vec / vec_per_field) * g.nz + pz0) * nxy;
      for (int i = tid; i < nx * nrows; i += nt) {
        const int ix = i % nx, iy = i / nx;
        plane[iy * pitch + ix] = cmul(f[i], r[i]);
      }
    } else {
      for (int i = tid; i < nx * nrows; i += nt) {
        const int ix = i % nx, iy = i / nx;
        plane[iy * pitch + ix] = r[i];
      }
    }
  }
  if (MODE == PLANE_FROM_R) {
    __syncthreads();
    // forward along x (permuted in -> natural out), all rows
    if (g.rx2 > 1) {
      run_contig<-1>(g.rx2, plane, nrows, nullptr, pitch, 1, g.rx1, twx, true, tid, nt);
      __syncthreads();
    }
    run_strided<-1>(g.rx1, plane, nrows, nullptr, pitch, 1, g.rx2, twx, false, tid, nt);
    __syncthreads();
  }
  // forward along y only for the x columns of the output sphere
  if (g.ry2 > 1) {
    run_contig<-1>(g.ry2, plane, np * sout.nxs, ids_out, 1, pitch, g.ry1, twy, true, tid, nt);
    __syncthreads();
  }
  run_strided<-1>(g.ry1, plane, np * sout.nxs, ids_out, 1, pitch, g.ry2, twy, false, tid, nt);
  __syncthreads();
  cplx *orow = Tout + ((long)vec * g.nz + pz0) * sout.ncol;
  for (int i = tid; i < np * sout.ncol; i += nt) {
    const int p = ONE ? 0 : i / sout.ncol, c = i - p * sout.ncol;
    orow[i] = plane[p * psz + sout.col_off[c]];
  }
}


// ------------------------------------------------------------------------------------------------ persistent VLOC plane kernel
// H.psi local part, one (vector, z-plane) item at a time on a PERSISTENT CTA (2 per SM), compile-time radix plan:
//   * the item's input row T[vec][pz][0..ncol) (one contiguous 16 B x ncol segment) is fetched by the TMA engine
//     (cp.async.bulk global -> shared, completion on an mbarrier) into a staging buffer while the previous item is still
//     being transformed: the fetch of item i+1 is issued as soon as the first stage of item i has consumed the buffer;
//   * the first inverse y stage reads its inputs straight from the staged row through a per-task index table (int16,
//     built with the sphere) and the last forward y stage writes its outputs straight to the output row in global memory:
//     no zero fill of the plane, no scatter / gather sweeps;
//   * the two strided x stages skip the x columns that hold no sphere data (loads of the inverse stage, stores of the
//     forward stage) with one bit mask per sub-index, uniform over (almost) every warp;
//   * twiddles, tables and masks are loaded once per CTA, not once per plane.
// Arithmetic per element is that of k_plane<PLANE_VLOC> (same codelets, same twiddle products, same order): results are
// bit-identical, which tests/test_gpu_operator.py checks against the generic kernel.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP); bytes must be a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// stage_mid (fft_core.h) with the potential stored y-fastest: v of (row l, x = a R + j) at vt[(a R + j) * vts + l].
// Same arithmetic; only the addresses of the 8-byte loads differ (coalesced over the rows a warp works on).
template <int R>
__device__ __forceinline__ void stage_mid_vt(cplx *x, int nlines, int ls, int r_other, const cplx *tw, const double *vt, int vts,
                                             int tid, int nthreads) {
  const int ntasks = nlines * r_other;
  TaskIter it(tid, nthreads, nlines);
  for (int t = tid; t < ntasks; t += nthreads, it.next()) {
    const int l = it.l, a = it.j;
    cplx *base = x + (l * ls + a * R);
    const double *vr = vt + (a * R) * vts + l;
    double q[R];
#pragma unroll
    for (int j = 0; j < R; ++j) q[j] = vr[j * vts];
    double re[R], im[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const cplx w = base[j]; re[j] = w.x; im[j] = w.y; }
    dft_fwd<R>(im, re);
#pragma unroll
    for (int j = 0; j < R; ++j) { re[j] *= q[j]; im[j] *= q[j]; }
    dft_fwd<R>(re, im);
#pragma unroll
    for (int k = 1; k < R; ++k) {
      const cplx w = tw[a * k];
      const double c = w.x, s = w.y;
      const double p = re[k], qq = im[k];
      re[k] = p * c - qq * s;
      im[k] = p * s + qq * c;
    }
#pragma unroll
    for (int k = 0; k < R; ++k) base[k] = cmake(re[k], im[k]);
  }
}

// stage_contig (fft_core.h) over x columns (line start = line_ids[l], element stride es) with the line count padded to a
// multiple of 8 in the task mapping (see the y stages of k_plane_vloc); same arithmetic
template <int R, int DIR>
__device__ __forceinline__ void stage_contig_pad(cplx *x, int nlines, const int *line_ids, int es, int r_other, const cplx *tw,
                                                 bool do_tw, int tid, int nthreads, int pad) {
  const int nlp = pad ? (nlines + 7) & ~7 : nlines;
  const int ntasks = nlp * r_other;
  TaskIter it(tid, nthreads, nlp);
  for (int t = tid; t < ntasks; t += nthreads, it.next()) {
    if (it.l >= nlines) continue;
    const int a = it.j;
    cplx *base = x + (line_ids[it.l] + a * R * es);
    double re[R], im[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const cplx v = base[j * es]; re[j] = v.x; im[j] = v.y; }
    if (DIR < 0) dft_fwd<R>(re, im); else dft_fwd<R>(im, re);
    if (do_tw) {
#pragma unroll
      for (int k = 1; k < R; ++k) {
        const cplx w = tw[a * k];
        const double c = w.x, s = DIR < 0 ? w.y : -w.y;
        const double p = re[k], q = im[k];
        re[k] = p * c - q * s;
        im[k] = p * s + q * c;
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) base[k * es] = cmake(re[k], im[k]);
  }
}

struct PlaneVArgs {
  const cplx *twx, *twy;
  const cplx *Tin;
  cplx *Tout;
  const double *vperm;      // y-fastest copy of the potential, [pz][px][py]
  const int *active;
  const int *xs;            // x columns that hold sphere data (ascending)
  const short *ytab;        // [(j2 * nxs + l) * RY1P + k] -> column index of (x = xs[l], y = j2 + RY2 k) in a T row, or -1
  int nz, nvec, ncol, nxs, ypad;
};

// named barrier over one thread group (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// The CTA is split into G groups of NT / G threads.  A group owns NY / G consecutive rows during the x transforms and a
// contiguous share of the data-holding x columns during the y transforms: the 1-D transforms of a line only touch that line,
// so the stages of one direction are separated by GROUP barriers (bar.sync id, NT / G) and the whole CTA meets only twice
// per item, where the direction changes (y -> x and x -> y).
template <int RX1, int RX2, int RY1, int RY2, int NT, int G, bool XMASK>
__global__ void __launch_bounds__(NT, 2) k_plane_vloc(PlaneVArgs a) {
  constexpr int NX = RX1 * RX2, NY = RY1 * RY2, PITCH = NX | 1;
  constexpr int GS = NT / G, ROWS = NY / G;
  constexpr int RY1P = (RY1 + 7) & ~7;        // table entries per task, padded to whole 16-byte loads
  static_assert(NT % G == 0 && GS % 32 == 0 && NY % G == 0, "group split");
  const int tid = threadIdx.x;
  const int grp = tid / GS, gt = tid - grp * GS;
  extern __shared__ __align__(128) unsigned char psm[];
  cplx *plane = (cplx *)psm;
  cplx *stage = plane + NY * PITCH;
  const int ncol_pad = (a.ncol + 7) & ~7;
  cplx *twx = stage + ncol_pad;
  cplx *twy = twx + NX;
  short *ytab = (short *)(twy + NY);
  const int nxs = a.nxs, nxz = NX - nxs;
  const int ntab = nxs * RY2 * RY1P;
  int *xs = (int *)(ytab + ntab);
  int *xz = xs + NX;                       // complement of xs (columns without data)
  unsigned *xmask = (unsigned *)(xz + NX); // per sub-index j2 < RX2: bit k set <=> column j2 + RX2 k holds data
  unsigned long long *bar = (unsigned long long *)(xmask + ((RX2 + 1) & ~1));
  for (int i = tid; i < NX; i += NT) twx[i] = a.twx[i];
  for (int i = tid; i < NY; i += NT) twy[i] = a.twy[i];
  for (int i = tid; i < ntab / 8; i += NT) ((uint4 *)ytab)[i] = ((const uint4 *)a.ytab)[i];
  for (int i = tid; i < nxs; i += NT) xs[i] = a.xs[i];
  if (tid == 0) {
    int q = 0;
    for (int x = 0; x < NX; ++x) {
      bool used = false;
      for (int i = 0; i < nxs; ++i) used |= (a.xs[i] == x);
      if (!used) xz[q++] = x;
    }
    for (int j2 = 0; j2 < RX2; ++j2) {
      unsigned m = 0;
      for (int k = 0; k < RX1; ++k) {
        const int x = j2 + RX2 * k;
        bool used = false;
        for (int i = 0; i < nxs; ++i) used |= (a.xs[i] == x);
        if (used) m |= 1u << k;
      }
      xmask[j2] = m;
    }
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  // this group's share of the data columns [c0, c0 + ncg) and rows [r0, r0 + ROWS)
  const int cper = (nxs + G - 1) / G;
  const int c0 = min(nxs, grp * cper), ncg = min(nxs, c0 + cper) - c0;
  const int r0 = grp * ROWS;
  const int bid = 1 + grp;

  const long nitems = (long)a.nvec * a.nz;
  const unsigned row_bytes = (unsigned)a.ncol * (unsigned)sizeof(cplx);
  auto next_active = [&](long it) {
    while (it < nitems && a.active && !a.active[it / a.nz]) it += gridDim.x;
    return it;
  };
  long item = next_active(blockIdx.x);
  if (tid == 0 && item < nitems) {
    mbar_expect_tx(bar, row_bytes);
    tma_load_1d(stage, a.Tin + item * a.ncol, row_bytes, bar);
  }
  unsigned parity = 0;
  const cplx zero = cmake(0.0, 0.0);
  while (item < nitems) {
    const int pz = (int)(item % a.nz);
    if (!XMASK) {
      // columns without sphere data: the x stages read them as zeros (rows of this group; read again after the y -> x barrier)
      for (int i = gt; i < nxz * ROWS; i += GS) plane[xz[i % nxz] + (r0 + i / nxz) * PITCH] = zero;
    }
    mbar_wait(bar, parity);
    parity ^= 1u;
    if (ncg > 0) {  // ---- inverse y, stage 1 (strided DFT_RY1 + twiddle), inputs from the staged row
      // lines are padded to a multiple of 8 per sub-index: every quarter-warp then works on 8 consecutive columns of ONE
      // sub-index, i.e. on 8 distinct 16-byte bank groups (the padding lanes idle)
      const int ncp = a.ypad ? (ncg + 7) & ~7 : ncg;
      const int ntask = ncp * RY2;
      TaskIter it(gt, GS, ncp);
      for (int t = gt; t < ntask; t += GS, it.next()) {
        if (it.l >= ncg) continue;
        const int l = c0 + it.l, j2 = it.j;
        short tb[RY1P];
#pragma unroll
        for (int q = 0; q < RY1P / 8; ++q) *(uint4 *)(tb + 8 * q) = *(const uint4 *)(ytab + (j2 * nxs + l) * RY1P + 8 * q);
        double re[RY1], im[RY1];
#pragma unroll
        for (int k = 0; k < RY1; ++k) {
          const int idx = tb[k];
          const cplx v = idx >= 0 ? stage[idx] : zero;
          re[k] = v.x; im[k] = v.y;
        }
        dft_fwd<RY1>(im, re);
#pragma unroll
        for (int k = 1; k < RY1; ++k) {
          const cplx w = twy[j2 * k];
          const double c = w.x, sn = -w.y;
          const double p = re[k], q = im[k];
          re[k] = p * c - q * sn;
          im[k] = p * sn + q * c;
        }
        cplx *base = plane + xs[l] + j2 * PITCH;
#pragma unroll
        for (int k = 0; k < RY1; ++k) base[k * (RY2 * PITCH)] = cmake(re[k], im[k]);
      }
    }
    group_sync(bid, GS);
    // ---- inverse y, stage 2 (contiguous DFT_RY2)
    stage_contig_pad<RY2, +1>(plane, ncg, xs + c0, PITCH, RY1, twy, false, gt, GS, a.ypad);
    __syncthreads();                                   // y -> x: every column is needed by every row
    // the staging buffer is free: fetch the next item's row while this one is transformed
    const long next = next_active(item + gridDim.x);
    if (tid == 0 && next < nitems) {
      mbar_expect_tx(bar, row_bytes);
      tma_load_1d(stage, a.Tin + next * a.ncol, row_bytes, bar);
    }
    cplx *rows = plane + r0 * PITCH;
    {  // ---- inverse x, stage 1 (strided DFT_RX1 + twiddle), rows of this group; columns without data are not read
      constexpr int ntask = ROWS * RX2;
      TaskIter it(gt, GS, ROWS);
      for (int t = gt; t < ntask; t += GS, it.next()) {
        const int j2 = it.j;
        cplx *base = rows + it.l * PITCH + j2;
        const unsigned m = XMASK ? xmask[j2] : 0xffffffffu;
        double re[RX1], im[RX1];
#pragma unroll
        for (int k = 0; k < RX1; ++k) {
          const cplx v = (m >> k) & 1u ? base[k * RX2] : zero;
          re[k] = v.x; im[k] = v.y;
        }
        dft_fwd<RX1>(im, re);
#pragma unroll
        for (int k = 1; k < RX1; ++k) {
          const cplx w = twx[j2 * k];
          const double c = w.x, sn = -w.y;
          const double p = re[k], q = im[k];
          re[k] = p * c - q * sn;
          im[k] = p * sn + q * c;
        }
#pragma unroll
        for (int k = 0; k < RX1; ++k) base[k * RX2] = cmake(re[k], im[k]);
      }
    }
    group_sync(bid, GS);
    // ---- last inverse x stage, x v(r), first forward x stage: one register round trip
    stage_mid_vt<RX2>(rows, ROWS, PITCH, RX1, twx, a.vperm + (long)pz * (NX * NY) + r0, NY, gt, GS);
    group_sync(bid, GS);
    {  // ---- forward x, last stage (strided DFT_RX1): only the columns of the output sphere are stored
      constexpr int ntask = ROWS * RX2;
      TaskIter it(gt, GS, ROWS);
      for (int t = gt; t < ntask; t += GS, it.next()) {
        const int j2 = it.j;
        cplx *base = rows + it.l * PITCH + j2;
        const unsigned m = XMASK ? xmask[j2] : 0xffffffffu;
        double re[RX1], im[RX1];
#pragma unroll
        for (int k = 0; k < RX1; ++k) { const cplx v = base[k * RX2]; re[k] = v.x; im[k] = v.y; }
        dft_fwd<RX1>(re, im);
#pragma unroll
        for (int k = 0; k < RX1; ++k)
          if ((m >> k) & 1u) base[k * RX2] = cmake(re[k], im[k]);
      }
    }
    __syncthreads();                                   // x -> y
    // ---- forward y, stage 1 (contiguous DFT_RY2 + twiddle)
    stage_contig_pad<RY2, -1>(plane, ncg, xs + c0, PITCH, RY1, twy, true, gt, GS, a.ypad);
    group_sync(bid, GS);
    if (ncg > 0) {  // ---- forward y, stage 2 (strided DFT_RY1): sphere entries go straight to the output row
      cplx *orow = a.Tout + item * a.ncol;
      const int ncp = a.ypad ? (ncg + 7) & ~7 : ncg;
      const int ntask = ncp * RY2;
      TaskIter it(gt, GS, ncp);
      for (int t = gt; t < ntask; t += GS, it.next()) {
        if (it.l >= ncg) continue;
        const int l = c0 + it.l, j2 = it.j;
        const cplx *base = plane + xs[l] + j2 * PITCH;
        double re[RY1], im[RY1];
#pragma unroll
        for (int k = 0; k < RY1; ++k) { const cplx v = base[k * (RY2 * PITCH)]; re[k] = v.x; im[k] = v.y; }
        dft_fwd<RY1>(re, im);
        short tb[RY1P];
#pragma unroll
        for (int q = 0; q < RY1P / 8; ++q) *(uint4 *)(tb + 8 * q) = *(const uint4 *)(ytab + (j2 * nxs + l) * RY1P + 8 * q);
#pragma unroll
        for (int k = 0; k < RY1; ++k) {
          const int idx = tb[k];
          if (idx >= 0) orow[idx] = cmake(re[k], im[k]);
        }
      }
    }
    group_sync(bid, GS);                               // the group's columns are rewritten by the next item's first stage
    item = next;
  }
}

template <int R>
__device__ __forceinline__ void stage_acc(const cplx *x, cplx *acc, int nlines, int ls, int r_other, const cplx *pr, int vls,
                                          int tid, int nthreads) {
  const int ntasks = nlines * r_other;
  TaskIter it(tid, nthreads, nlines);
  for (int t = tid; t < ntasks; t += nthreads, it.next()) {
    const int l = it.l, a = it.j;
    const cplx *base = x + (l * ls + a * R);
    double re[R], im[R];
#pragma unroll
    for (int j = 0; j < R; ++j) { const cplx w = base[j]; re[j] = w.x; im[j] = w.y; }
    dft_fwd<R>(im, re);
    cplx *ab = acc + (l * ls + a * R);
    const cplx *pb = pr + (l * vls + a * R);
#pragma unroll
    for (int j = 0; j < R; ++j) ab[j] = cfma(cconj(pb[j]), cmake(re[j], im[j]), ab[j]);   // drho += conj(psi) dpsi
  }
}
__device__ __noinline__ void run_acc_big(int R, const cplx *x, cplx *acc, int nlines, int ls, int r_other, const cplx *pr, int vls,
                                         int tid, int nthreads) {
  switch (R) {
#define SGW_CASE(r) case r: stage_acc<r>(x, acc, nlines, ls, r_other, pr, vls, tid, nthreads); break;
    SGW_FOR_EACH_BIG_RADIX(SGW_CASE)
#undef SGW_CASE
    default: break;
  }
}
__device__ __noinline__ void run_acc(int R, const cplx *x, cplx *acc, int nlines, int ls, int r_other, const cplx *pr, int vls,
                                     int tid, int nthreads) {
  switch (R) {
#define SGW_CASE(r) case r: stage_acc<r>(x, acc, nlines, ls, r_other, pr, vls, tid, nthreads); break;
    SGW_FOR_EACH_RADIX(SGW_CASE)
#undef SGW_CASE
    default: run_acc_big(R, x, acc, nlines, ls, r_other, pr, vls, tid, nthreads); break;
  }
}

// incdrhoscf: one CTA per (pf = perturbation x frequency, z-plane); bands are summed on chip
template <int NT, bool GMEM = false>
__global__ void __launch_bounds__(NT, NT >= 512 ? 1 : (NT >= 256 ? 2 : 3)) k_plane_rho(GridDev g, SphereDev sin, SphereDev sout, int nocc,
                                                         const cplx *__restrict__ Tin, const cplx *__restrict__ psir,
                                                         double wgt, cplx *__restrict__ Tout, int accumulate,
                                                         cplx *scratch = nullptr, int pf0 = 0) {
  const int pf = blockIdx.x + (GMEM ? pf0 : 0), pz = blockIdx.y;   // pf fastest: concurrent CTAs share the psi_v(r) planes through L2
  const int tid = threadIdx.x, nt = blockDim.x;
  extern __shared__ cplx sm[];
  const int pitch = g.pitchx, nx = g.nx, ny = g.ny;
  // GMEM: plane and accumulator of this CTA in global memory (boxes whose planes do not fit in shared memory, see k_plane)
  cplx *plane = GMEM ? scratch + ((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * ny * pitch : sm;
  cplx *acc = plane + ny * pitch;
  cplx *twx = GMEM ? sm : acc + ny * pitch;
  cplx *twy = twx + nx;
  int *xs_in = (int *)(twy + ny), *xs_out = xs_in + nx;
  for (int i = tid; i < nx; i += nt) twx[i] = g.twx[i];
  for (int i = tid; i < ny; i += nt) twy[i] = g.twy[i];
  for (int i = tid; i < sin.nxs; i += nt) xs_in[i] = sin.xs[i];
  for (int i = tid; i < sout.nxs; i += nt) xs_out[i] = sout.xs[i];
  for (int i = tid; i < ny * pitch; i += nt) acc[i] = cmake(0.0, 0.0);
  const long nxy = (long)nx * ny;
  for (int ib = 0; ib < nocc; ++ib) {
    for (int i = tid; i < ny * pitch; i += nt) plane[i] = cmake(0.0, 0.0);
    __syncthreads();
    const cplx *row = Tin + (((long)pf * nocc + ib) * g.nz + pz) * sin.ncol;
    for (int c = tid; c < sin.ncol; c += nt) plane[sin.col_off[c]] = row[c];
    __syncthreads();
    run_strided<+1>(g.ry1, plane, sin.nxs, xs_in, 1, pitch, g.ry2, twy, g.ry2 > 1, tid, nt);
    __syncthreads();
    if (g.ry2 > 1) {
      run_contig<+1>(g.ry2, plane, sin.nxs, xs_in, 1, pitch, g.ry1, twy, false, tid, nt);
      __syncthreads();
    }
    const cplx *pr = psir + ((long)ib * g.nz + pz) * nxy;
    if (g.rx2 > 1) {
      run_strided<+1>(g.rx1, plane, ny, nullptr, pitch, 1, g.rx2, twx, true, tid, nt);
      __syncthreads();
      run_acc(g.rx2, plane, acc, ny, pitch, g.rx1, pr, nx, tid, nt);
    } else {
      run_acc(g.rx1, plane, acc, ny, pitch, 1, pr, nx, tid, nt);
    }
    __syncthreads();
  }
  for (int i = tid; i < nx * ny; i += nt) {
    const int ix = i % nx, iy = i / nx;
    const int j = iy * pitch + ix;
    plane[j] = cscale(wgt, acc[j]);
  }
  __syncthreads();
  if (g.rx2 > 1) {
    run_contig<-1>(g.rx2, plane, ny, nullptr, pitch, 1, g.rx1, twx, true, tid, nt);
    __syncthreads();
  }
  run_strided<-1>(g.rx1, plane, ny, nullptr, pitch, 1, g.rx2, twx, false, tid, nt);
  __syncthreads();
  if (g.ry2 > 1) {
    run_contig<-1>(g.ry2, plane, sout.nxs, xs_out, 1, pitch, g.ry1, twy, true, tid, nt);
    __syncthreads();
  }
  run_strided<-1>(g.ry1, plane, sout.nxs, xs_out, 1, pitch, g.ry2, twy, false, tid, nt);
  __syncthreads();
  cplx *orow = Tout + ((long)pf * g.nz + pz) * sout.ncol;
  if (accumulate) {
    for (int c = tid; c < sout.ncol; c += nt) orow[c] = cadd(orow[c], plane[sout.col_off[c]]);
  } else {
    for (int c = tid; c < sout.ncol; c += nt) orow[c] = plane[sout.col_off[c]];
  }
}



// ------------------------------------------------------------------------------------------------ Delta-rho plane kernel, v2
// incdrhoscf for one (perturbation x frequency, z-plane) per CTA, compile-time radix plan, the structure of k_plane_vloc:
//   * the band's input row arrives by cp.async.bulk + mbarrier, the next band's row is in flight while this one is transformed;
//   * the first inverse y stage gathers from the staged row through the index table (no zero fill, no scatter), the strided x
//     stage skips empty columns;
//   * the last inverse x stage is fused with the accumulation  acc += conj(psi_v(r)) dpsi_v(r)  and the accumulator lives in
//     REGISTERS: with NT >= NY * RX1 every thread owns at most one (row, x group) task, the same one for every band, so the
//     per-band read-modify-write of an accumulator plane in shared memory disappears;
//   * psi_v(r) is read from a copy stored y-fastest, so consecutive lanes (consecutive rows) read consecutive addresses.
// Same butterflies and the same band order as k_plane_rho: bit-identical result (GPU test).
struct PlaneRhoArgs {
  const cplx *twx, *twy;
  const cplx *Tin;           // [(pf * nocc + ib) * nz + pz][ncol_in]
  const cplx *psir_t;        // [ib][pz][x][y]
  cplx *Tout;                // [pf * nz + pz][ncol_out]
  const int *xs_in, *xs_out; // data-holding x columns of the two spheres
  const short *ytab_in;      // gather table of the input sphere (build_ytab)
  const int *col_off_out;    // plane offset of every column of the output sphere
  int nz, nocc, ncol_in, nxs_in, ncol_out, nxs_out, accumulate;
  double wgt;
};

template <int RX1, int RX2, int RY1, int RY2, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_plane_rho_v2(PlaneRhoArgs a) {
  constexpr int NX = RX1 * RX2, NY = RY1 * RY2, PITCH = NX | 1;
  constexpr int RY1P = (RY1 + 7) & ~7;
  static_assert(NT >= NY * RX1, "one accumulation task per thread");
  const int pf = blockIdx.x, pz = blockIdx.y;   // pf fastest: concurrent CTAs share the psi_v(r) planes through L2
  const int tid = threadIdx.x;
  extern __shared__ __align__(128) unsigned char psm[];
  cplx *plane = (cplx *)psm;
  cplx *stage = plane + NY * PITCH;
  const int ncol_pad = (a.ncol_in + 7) & ~7;
  cplx *twx = stage + ncol_pad;
  cplx *twy = twx + NX;
  short *ytab = (short *)(twy + NY);
  const int nxs = a.nxs_in;
  const int ntab = nxs * RY2 * RY1P;
  int *xs = (int *)(ytab + ntab);
  int *xso = xs + NX;
  unsigned *xmask = (unsigned *)(xso + NX);
  unsigned long long *bar = (unsigned long long *)(xmask + ((RX2 + 1) & ~1));
  for (int i = tid; i < NX; i += NT) twx[i] = a.twx[i];
  for (int i = tid; i < NY; i += NT) twy[i] = a.twy[i];
  for (int i = tid; i < ntab / 8; i += NT) ((uint4 *)ytab)[i] = ((const uint4 *)a.ytab_in)[i];
  for (int i = tid; i < nxs; i += NT) xs[i] = a.xs_in[i];
  for (int i = tid; i < a.nxs_out; i += NT) xso[i] = a.xs_out[i];
  if (tid < RX2) {
    unsigned m = 0;
    for (int k = 0; k < RX1; ++k) {
      const int x = tid + RX2 * k;
      bool used = false;
      for (int i = 0; i < nxs; ++i) used |= (a.xs_in[i] == x);
      if (used) m |= 1u << k;
    }
    xmask[tid] = m;
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  __syncthreads();
  const unsigned row_bytes = (unsigned)a.ncol_in * (unsigned)sizeof(cplx);
  const cplx *rows = a.Tin + ((long)pf * a.nocc * a.nz + pz) * a.ncol_in;     // band ib at + ib * nz * ncol_in
  const long band_stride = (long)a.nz * a.ncol_in;
  if (tid == 0) {
    mbar_expect_tx(bar, row_bytes);
    tma_load_1d(stage, rows, row_bytes, bar);
  }
  // the accumulation task of this thread: row l, x group g (contiguous DFT_RX2 over x = g RX2 .. g RX2 + RX2 - 1)
  const bool has_acc = tid < NY * RX1;
  const int al = tid % NY, ag = tid / NY;
  double accr[RX2], acci[RX2];
#pragma unroll
  for (int j = 0; j < RX2; ++j) accr[j] = acci[j] = 0.0;
  unsigned parity = 0;
  const cplx zero = cmake(0.0, 0.0);
  for (int ib = 0; ib < a.nocc; ++ib) {
    mbar_wait(bar, parity);
    parity ^= 1u;
    {  // inverse y, stage 1 (strided DFT_RY1 + twiddle) from the staged row
      const int ntask = nxs * RY2;
      TaskIter it(tid, NT, nxs);
      for (int t = tid; t < ntask; t += NT, it.next()) {
        const int l = it.l, j2 = it.j;
        short tb[RY1P];
#pragma unroll
        for (int q = 0; q < RY1P / 8; ++q) *(uint4 *)(tb + 8 * q) = *(const uint4 *)(ytab + (j2 * nxs + l) * RY1P + 8 * q);
        double re[RY1], im[RY1];
#pragma unroll
        for (int k = 0; k < RY1; ++k) {
          const int idx = tb[k];
          const cplx v = idx >= 0 ? stage[idx] : zero;
          re[k] = v.x; im[k] = v.y;
        }
        dft_fwd<RY1>(im, re);
#pragma unroll
        for (int k = 1; k < RY1; ++k) {
          const cplx w = twy[j2 * k];
          const double c = w.x, sn = -w.y;
          const double p = re[k], q = im[k];
          re[k] = p * c - q * sn;
          im[k] = p * sn + q * c;
        }
        cplx *base = plane + xs[l] + j2 * PITCH;
#pragma unroll
        for (int k = 0; k < RY1; ++k) base[k * (RY2 * PITCH)] = cmake(re[k], im[k]);
      }
    }
    __syncthreads();
    if (tid == 0 && ib + 1 < a.nocc) {          // the staging buffer is free: fetch the next band's row
      mbar_expect_tx(bar, row_bytes);
      tma_load_1d(stage, rows + (long)(ib + 1) * band_stride, row_bytes, bar);
    }
    stage_contig<RY2, +1>(plane, nxs, xs, 1, PITCH, RY1, twy, false, tid, NT);
    __syncthreads();
    {  // inverse x, stage 1 (strided DFT_RX1 + twiddle), all rows; columns without data are read as zeros
      const int ntask = NY * RX2;
      TaskIter it(tid, NT, NY);
      for (int t = tid; t < ntask; t += NT, it.next()) {
        const int j2 = it.j;
        cplx *base = plane + it.l * PITCH + j2;
        const unsigned m = xmask[j2];
        double re[RX1], im[RX1];
#pragma unroll
        for (int k = 0; k < RX1; ++k) {
          const cplx v = (m >> k) & 1u ? base[k * RX2] : zero;
          re[k] = v.x; im[k] = v.y;
        }
        dft_fwd<RX1>(im, re);
#pragma unroll
        for (int k = 1; k < RX1; ++k) {
          const cplx w = twx[j2 * k];
          const double c = w.x, sn = -w.y;
          const double p = re[k], q = im[k];
          re[k] = p * c - q * sn;
          im[k] = p * sn + q * c;
        }
#pragma unroll
        for (int k = 0; k < RX1; ++k) base[k * RX2] = cmake(re[k], im[k]);
      }
    }
    __syncthreads();
    if (has_acc) {  // last inverse x stage (contiguous DFT_RX2) fused with acc += conj(psi_v(r)) dpsi_v(r), accumulator in registers
      const cplx *base = plane + al * PITCH + ag * RX2;
      const cplx *pr = a.psir_t + (((long)ib * a.nz + pz) * NX + ag * RX2) * NY + al;
      cplx pv[RX2];
#pragma unroll
      for (int j = 0; j < RX2; ++j) pv[j] = pr[(long)j * NY];
      double re[RX2], im[RX2];
#pragma unroll
      for (int j = 0; j < RX2; ++j) { const cplx w = base[j]; re[j] = w.x; im[j] = w.y; }
      dft_fwd<RX2>(im, re);
#pragma unroll
      for (int j = 0; j < RX2; ++j) {
        const cplx r = cfma(cconj(pv[j]), cmake(re[j], im[j]), cmake(accr[j], acci[j]));   // drho += conj(psi) dpsi
        accr[j] = r.x; acci[j] = r.y;
      }
    }
    __syncthreads();                             // the next band's first stage rewrites the data columns
  }
  if (has_acc) {
    cplx *base = plane + al * PITCH + ag * RX2;
#pragma unroll
    for (int j = 0; j < RX2; ++j) base[j] = cscale(a.wgt, cmake(accr[j], acci[j]));
  }
  __syncthreads();
  // forward 2-D transform of the sum, only the columns of the density sphere; then its columns -> Tout
  stage_contig<RX2, -1>(plane, NY, nullptr, PITCH, 1, RX1, twx, true, tid, NT);
  __syncthreads();
  stage_strided<RX1, -1>(plane, NY, nullptr, PITCH, 1, RX2, twx, false, tid, NT);
  __syncthreads();
  stage_contig<RY2, -1>(plane, a.nxs_out, xso, 1, PITCH, RY1, twy, true, tid, NT);
  __syncthreads();
  stage_strided<RY1, -1>(plane, a.nxs_out, xso, 1, PITCH, RY2, twy, false, tid, NT);
  __syncthreads();
  cplx *orow = a.Tout + ((long)pf * a.nz + pz) * a.ncol_out;
  if (a.accumulate) {
    for (int c = tid; c < a.ncol_out; c += NT) orow[c] = cadd(orow[c], plane[a.col_off_out[c]]);
  } else {
    for (int c = tid; c < a.ncol_out; c += NT) orow[c] = plane[a.col_off_out[c]];
  }
}

// out[p][x][y] = in[p][y][x] for nplanes planes (the y-fastest copy of psi_v(r) the accumulation stage reads)
__global__ void k_transpose_planes(int nx, int ny, const cplx *__restrict__ in, cplx *__restrict__ out) {
  __shared__ cplx tile[16][17];
  const long p = blockIdx.z;
  const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (x0 + tx < nx && y0 + ty < ny) tile[ty][tx] = in[(p * ny + y0 + ty) * nx + x0 + tx];
  __syncthreads();
  if (x0 + ty < nx && y0 + tx < ny) out[(p * nx + x0 + ty) * ny + y0 + tx] = tile[tx][ty];
}

// ------------------------------------------------------------------------------------------------ persistent z passes (TMA)
// The z transforms are pure HBM streams (2 x 16 B x nz x ncol + the vector per H.psi).  One CTA owns ONE block of ZB = 16
// sphere columns (its index tables, twiddles and kinetic energies are loaded once) and walks over the vectors of the batch:
//   g2r: the block's entries of vector v+1 -- one contiguous range of the column-ordered vector -- are fetched by a 1-D bulk
//        copy (cp.async.bulk + mbarrier) while vector v is transformed; the first stage gathers its inputs from that staged
//        range through an index table (no zero fill, no scatter pass); the finished [nz][16] tile leaves through ONE 3-D
//        tensor-map store (cp.async.bulk.tensor, SASS UTMASTG) that overlaps the next vector's work (two tile buffers,
//        cp.async.bulk.wait_group.read before a buffer is reused);
//   r2g: the [nz][16] tile of vector v+1 is fetched by a tensor-map load (UTMALDG) into the other buffer while vector v is
//        transformed; psi / the non-local part of the epilogue are prefetched into registers before the wait.
// Tile layout [z][c]: element stride 16, line stride 1 -- the lanes of a quarter-warp work on 8 consecutive columns, i.e. on 8
// distinct 16-byte bank groups, and the tile is exactly the dense box the tensor map describes.
constexpr int ZB = 16;
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d(const void *tmap, const void *smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(tmap), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const void *tmap, int c0, int c1, int c2, unsigned long long *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }

struct ZTArgs {
  const cplx *twz;
  const int *col_ptr;
  const short *ztab;        // [(c * RZ2 + j2) * RZ1P + k] -> offset of (column c, z = j2 + RZ2 k) from the column's first entry, or -1
  const cplx *vec;          // g2r: input vectors ; r2g: psi of the epilogue
  cplx *out;                // r2g: output vectors
  long ld;
  const int *active;
  int ncol, nvec, nz, maxlen;   // maxlen: largest number of entries of a column block, rounded up to a multiple of 8
  // r2g epilogue (ZEpilogue)
  int mode, keep_out;
  const double *g2kin;
  const cplx *sigma;
  long sigma_stride;
  double scale;
};

// DB: two staging buffers for the input ranges, so that the bulk copy of the vector after next is in flight during a whole
// iteration (with one buffer the copy can only be issued after stage 1 has consumed it and its latency is exposed: ncu showed
// long_scoreboard 53 % in the mbarrier wait)
// DBT: two output tiles, so that the tensor-map store of vector v may still be reading its tile while vector v + 1 is
// transformed into the other one (cp.async.bulk.wait_group.read 1 instead of 0)
template <int RZ1, int RZ2, int NT, bool DB, bool DBT = false>
__global__ void __launch_bounds__(NT, DBT ? 4 : 6) k_zpass_g2r_tma(const __grid_constant__ CUtensorMap tmap, ZTArgs a) {
  constexpr int NZ = RZ1 * RZ2, RZ1P = (RZ1 + 7) & ~7, TILE = NZ * ZB;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * ZB, nc = min(ZB, a.ncol - c0);
  extern __shared__ __align__(128) unsigned char zsm[];
  cplx *tile0 = (cplx *)zsm;                      // [NZ][ZB] (x 2 with DBT)
  cplx *in0 = tile0 + (DBT ? 2 : 1) * TILE;       // [maxlen] staged entries of the block (x 2 with DB)
  cplx *in1 = DB ? in0 + a.maxlen : in0;
  cplx *tw = in1 + a.maxlen;
  short *ztab = (short *)(tw + NZ);               // [ZB][RZ2][RZ1P], rebased to the block's first entry
  unsigned long long *bar = (unsigned long long *)(ztab + ZB * RZ2 * RZ1P);   // [2]
  const int p0 = a.col_ptr[c0], p1 = a.col_ptr[c0 + nc];
  const unsigned bytes = (unsigned)(p1 - p0) * (unsigned)sizeof(cplx);
  for (int i = tid; i < NZ; i += NT) tw[i] = a.twz[i];
  for (int i = tid; i < ZB * RZ2 * RZ1P; i += NT) {
    const int c = i / (RZ2 * RZ1P);
    short v = -1;
    if (c < nc) {
      v = a.ztab[(long)(c0 + c) * (RZ2 * RZ1P) + (i - c * (RZ2 * RZ1P))];
      if (v >= 0) v = (short)(v + a.col_ptr[c0 + c] - p0);
    }
    ztab[i] = v;
  }
  if (tid == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); fence_mbar_init(); }
  __syncthreads();
  auto next_active = [&](int v) {
    while (v < a.nvec && a.active && !a.active[v]) v += gridDim.y;
    return v;
  };
  int v = next_active(blockIdx.y);
  int v1 = v < a.nvec ? next_active(v + gridDim.y) : a.nvec;
  if (tid == 0 && bytes) {
    if (v < a.nvec) { mbar_expect_tx(bar, bytes); tma_load_1d(in0, a.vec + (long)v * a.ld + p0, bytes, bar); }
    if (DB && v1 < a.nvec) { mbar_expect_tx(bar + 1, bytes); tma_load_1d(in1, a.vec + (long)v1 * a.ld + p0, bytes, bar + 1); }
  }
  unsigned parity0 = 0u, parity1 = 0u;
  int sb = 0;                                     // staging buffer of the current vector
  int tb_sel = 0;                                 // output tile of the current vector (DBT)
  const cplx zero = cmake(0.0, 0.0);
  while (v < a.nvec) {
    const int vn = v1;
    const int vnn = vn < a.nvec ? next_active(vn + gridDim.y) : a.nvec;
    cplx *tile = tile0 + (DBT && tb_sel ? TILE : 0);
    if (tid == 0) {                               // the store that last used THIS tile has finished reading it
      if (DBT) bulk_wait_read<1>(); else bulk_wait_read<0>();
    }
    __syncthreads();
    const cplx *in = (DB && sb) ? in1 : in0;
    if (bytes) {
      if (DB && sb) { mbar_wait(bar + 1, parity1); parity1 ^= 1u; }
      else { mbar_wait(bar, parity0); parity0 ^= 1u; }
    }
    {  // inverse z, stage 1 (strided DFT_RZ1 + twiddle): inputs gathered from the staged range; all ZB columns are written
      constexpr int ntask = ZB * RZ2;
      for (int task = tid; task < ntask; task += NT) {
        const int c = task % ZB, j2 = task / ZB;
        short tb[RZ1P];
#pragma unroll
        for (int q = 0; q < RZ1P / 8; ++q) *(uint4 *)(tb + 8 * q) = *(const uint4 *)(ztab + (c * RZ2 + j2) * RZ1P + 8 * q);
        double re[RZ1], im[RZ1];
#pragma unroll
        for (int k = 0; k < RZ1; ++k) {
          const int idx = tb[k];
          const cplx x = idx >= 0 ? in[idx] : zero;
          re[k] = x.x; im[k] = x.y;
        }
        dft_fwd<RZ1>(im, re);
#pragma unroll
        for (int k = 1; k < RZ1; ++k) {
          const cplx w = tw[j2 * k];
          const double cs = w.x, sn = -w.y;
          const double p = re[k], q = im[k];
          re[k] = p * cs - q * sn;
          im[k] = p * sn + q * cs;
        }
        cplx *base = tile + j2 * ZB + c;
#pragma unroll
        for (int k = 0; k < RZ1; ++k) base[k * (RZ2 * ZB)] = cmake(re[k], im[k]);
      }
    }
    __syncthreads();
    if (tid == 0 && bytes) {                      // this staging buffer is free: fetch the entries of the next vector that uses it
      if (DB) {
        if (vnn < a.nvec) {
          unsigned long long *bb = sb ? bar + 1 : bar;
          mbar_expect_tx(bb, bytes);
          tma_load_1d(sb ? in1 : in0, a.vec + (long)vnn * a.ld + p0, bytes, bb);
        }
      } else if (vn < a.nvec) {
        mbar_expect_tx(bar, bytes);
        tma_load_1d(in0, a.vec + (long)vn * a.ld + p0, bytes, bar);
      }
    }
    stage_contig<RZ2, +1>(tile, ZB, nullptr, 1, ZB, RZ1, tw, false, tid, NT);
    fence_proxy_async();                          // generic-proxy writes of the tile -> visible to the TMA engine
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&tmap, tile, 2 * c0, 0, v);
      bulk_commit();
    }
    v = vn;
    v1 = vnn;
    if (DB) sb ^= 1;
    if (DBT) tb_sel ^= 1;
  }
  if (tid == 0) bulk_wait_read<0>();              // shared memory must stay alive until the last store has read it
}

template <int RZ1, int RZ2, int NT, int EPT>
__global__ void __launch_bounds__(NT, 6) k_zpass_r2g_tma(const __grid_constant__ CUtensorMap tmap, ZTArgs a) {
  constexpr int NZ = RZ1 * RZ2, TILE = NZ * ZB;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * ZB, nc = min(ZB, a.ncol - c0);
  extern __shared__ __align__(128) unsigned char zsm[];
  cplx *tile = (cplx *)zsm;                       // [NZ][ZB]
  cplx *tw = tile + TILE;
  double *g2 = (double *)(tw + NZ);               // [maxlen] kinetic energies of the block's entries
  short *pos = (short *)(g2 + a.maxlen);          // [maxlen] z * ZB + c of every entry
  unsigned long long *bar = (unsigned long long *)(pos + a.maxlen);
  const int p0 = a.col_ptr[c0], p1 = a.col_ptr[c0 + nc], len = p1 - p0;
  for (int i = tid; i < NZ; i += NT) tw[i] = a.twz[i];
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  // entry -> tile position, from the gather table of the forward pass (same table: offset of (c, z) inside column c)
  for (int i = tid; i < nc * NZ; i += NT) {
    const int c = i / NZ, z = i - c * NZ;
    const int j2 = z % RZ2, k = z / RZ2;
    const short off = a.ztab[((long)(c0 + c) * RZ2 + j2) * ((RZ1 + 7) & ~7) + k];
    if (off >= 0) pos[off + a.col_ptr[c0 + c] - p0] = (short)(z * ZB + c);
  }
  if (a.mode == 1)
    for (int i = tid; i < len; i += NT) g2[i] = a.g2kin[p0 + i];
  __syncthreads();
  auto next_active = [&](int v) {
    while (v < a.nvec && a.active && !a.active[v]) v += gridDim.y;
    return v;
  };
  constexpr unsigned tile_bytes = TILE * sizeof(cplx);
  int v = next_active(blockIdx.y);
  if (tid == 0 && v < a.nvec) {
    mbar_expect_tx(bar, tile_bytes);
    tma_load_3d(tile, &tmap, 2 * c0, 0, v, bar);
  }
  unsigned parity = 0u;
  while (v < a.nvec) {
    const int vn = next_active(v + gridDim.y);
    cplx sg = cmake(0.0, 0.0);
    cplx *dst = a.out + (long)v * a.ld + p0;
    const cplx *psi = a.vec + (long)v * a.ld + p0;
    if (a.mode == 1 && a.sigma) sg = a.sigma[(long)v * a.sigma_stride];
    mbar_wait(bar, parity);
    parity ^= 1u;
    stage_contig<RZ2, -1>(tile, ZB, nullptr, 1, ZB, RZ1, tw, true, tid, NT);
    __syncthreads();
    stage_strided<RZ1, -1>(tile, ZB, nullptr, 1, ZB, RZ2, tw, false, tid, NT);
    __syncthreads();
    cplx res[EPT];
#pragma unroll
    for (int q = 0; q < EPT; ++q) {
      const int i = tid + q * NT;
      if (i < len) res[q] = tile[pos[i]];
    }
    fence_proxy_async();                          // this thread's generic-proxy accesses of the tile are ordered before the TMA write
    __syncthreads();                              // the tile has been read: fetch the next vector's tile while the epilogue runs
    if (tid == 0 && vn < a.nvec) {
      mbar_expect_tx(bar, tile_bytes);
      tma_load_3d(tile, &tmap, 2 * c0, 0, vn, bar);
    }
#pragma unroll
    for (int q = 0; q < EPT; ++q) {
      const int i = tid + q * NT;
      if (i < len) {
        cplx val = cscale(a.scale, res[q]);
        if (a.mode == 0) {
          dst[i] = val;
        } else if (a.mode == 1) {
          // (H + sigma) psi = g2kin psi + V_loc psi + [non-local, already in out] + sigma psi   (order of k_zpass_r2g)
          const cplx ps = psi[i];
          const cplx nl = a.keep_out ? dst[i] : cmake(0.0, 0.0);
          val = cadd(cscale(g2[i], ps), val);
          val = cadd(val, nl);
          dst[i] = cfma(sg, ps, val);
        } else {
          dst[i] = cadd(dst[i], val);
        }
      }
    }
    v = vn;
  }
}

// ------------------------------------------------------------------------------------------------ host side
GridDev grid_dev(const sgw_ctx *ctx, const FftGrid *gr) {
  GridDev g;
  if (gr) {
    g.nx = gr->n1; g.ny = gr->n2; g.nz = gr->n3;
    g.rx1 = gr->px.r1; g.rx2 = gr->px.r2;
    g.ry1 = gr->py.r1; g.ry2 = gr->py.r2;
    g.rz1 = gr->pz.r1; g.rz2 = gr->pz.r2;
    g.twx = gr->d_twx; g.twy = gr->d_twy; g.twz = gr->d_twz;
  } else {
    g.nx = ctx->nr1; g.ny = ctx->nr2; g.nz = ctx->nr3;
    g.rx1 = ctx->px.r1; g.rx2 = ctx->px.r2;
    g.ry1 = ctx->py.r1; g.ry2 = ctx->py.r2;
    g.rz1 = ctx->pz.r1; g.rz2 = ctx->pz.r2;
    g.twx = ctx->d_twx; g.twy = ctx->d_twy; g.twz = ctx->d_twz;
  }
  g.pitchx = g.nx | 1;
  return g;
}

static int zpass_zcb() {
  static int forced = -1;                                       // SGW_ZCB: tuning knob (8 | 16 | 32)
  if (forced < 0) { const char *e = getenv("SGW_ZCB"); forced = e ? atoi(e) : 0; }
  return (forced == 8 || forced == 32) ? forced : ZCB_DEFAULT;
}
static int zpass_threads(const sgw_ctx *ctx) {
  static int forced = -1;                                       // SGW_ZTHREADS: tuning knob (multiple of 32, <= ZTHREADS)
  if (forced < 0) { const char *e = getenv("SGW_ZTHREADS"); forced = e ? atoi(e) : 0; }
  if (forced >= 32 && forced <= (zpass_zcb() >= 32 ? 256 : 160)) return forced;
  (void)ctx;
  return 128;  // measured on B200 (Si64, 64-register build, >= 6 CTAs per SM): 96 -> 51.6 ms, 128 -> 49.6 ms, 160 -> 56.5 ms per step
}
static int plane_threads() {
  static int forced = -1;                                       // SGW_PTHREADS: tuning knob (256 | 384 | 512), default 384
  if (forced < 0) { const char *e = getenv("SGW_PTHREADS"); forced = e ? atoi(e) : 0; }
  return (forced == 256 || forced == 512) ? forced : 384;   // B200, Si64: 384 -> 155 ms, 256 -> 161 ms, 512 -> 162 ms per step
}
static size_t zpass_smem(const GridDev &g, int zcb) { return (size_t)(zcb * (g.nz | 1) + g.nz) * sizeof(cplx); }
static size_t plane_smem(const GridDev &g, int nplanes) {
  return (size_t)(nplanes * g.ny * g.pitchx + g.nx + g.ny) * sizeof(cplx) + 2 * (size_t)nplanes * g.nx * sizeof(int);
}
// z-planes per CTA of k_plane: enough points (~4096) for every thread of the radix stages, at most 8 planes
static int planes_per_cta(const GridDev &g) {
  static int forced = -1;                                       // SGW_PPC: tuning knob (1..8)
  if (forced < 0) { const char *e = getenv("SGW_PPC"); forced = e ? atoi(e) : 0; }
  int p = forced >= 1 && forced <= 8 ? forced : std::max(1, std::min(8, 4096 / (g.nx * g.ny)));
  return std::min(p, g.nz);
}

template <typename K>
static int set_smem(sgw_ctx *ctx, K kernel, size_t bytes) {
  if (bytes > ctx->smem_optin) {
    ctx->err = "FFT plane does not fit in shared memory";
    return SGW_E_UNSUPPORTED;
  }
  if (bytes > 48 * 1024) SGW_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SGW_OK;
}


// ---- host side of the TMA z passes
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_tiled_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}
// T[vec][pz][col] (complex FP64) as a rank-3 tensor of doubles {2 ncol, nz, nvec}; box = {2 ZB, nz, 1}: the [nz][ZB] tile of one vector
static bool make_tmap_T(CUtensorMap *tm, const cplx *T, int ncol, int nz, int nvec) {
  PFN_encodeTiled enc = encode_tiled_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)2 * ncol, (cuuint64_t)nz, (cuuint64_t)nvec};
  const cuuint64_t strides[2] = {(cuuint64_t)ncol * 16, (cuuint64_t)ncol * 16 * nz};
  const cuuint32_t box[3] = {2 * ZB, (cuuint32_t)nz, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)T, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// SGW_ZPASS: 0 = generic kernels, 1 = persistent TMA kernels for both passes, 2 = TMA for G -> r only (default), 3 = r -> G only.
// Measured at Si64, 1024 vectors per launch (us per launch, G -> r / r -> G): generic 475 / 510, TMA 447 / 610 -- the tensor-map
// LOAD of [nz][16] boxes (256-byte rows, 16 KB apart) is slower than the per-thread loads of the generic kernel, the bulk-copy
// staged input + tensor-map STORE of the forward pass is faster.
static int zpass_variant() {
  const char *e = getenv("SGW_ZPASS");
  return e ? atoi(e) : 2;
}
static bool zpass_tma_ok(const GridDev &g, const Sphere &s) {
  if (zpass_variant() == 0 || !s.d_ztab || s.ztab_rz1 != g.rz1 || s.ztab_rz2 != g.rz2 || g.nz > 256 || s.ncol < 1) return false;
  return (g.rz1 == 8 && g.rz2 == 9) || (g.rz1 == 5 && g.rz2 == 9);
}
template <int RZ1, int RZ2>
static int launch_zpass_tma(sgw_ctx *ctx, bool g2r, const CUtensorMap &tm, const ZTArgs &a) {
  constexpr int NZ = RZ1 * RZ2, NT = 128;
  const int ncb = (a.ncol + ZB - 1) / ZB;
  // SGW_ZG2R_DB=1: two input staging buffers.  Measured at Si64 (H.psi z passes per step): 36.1 ms with, 35.6 ms without -- the
  // bulk copy's latency is not what limits the kernel, so the default keeps one buffer (less shared memory per CTA)
  const char *edb = getenv("SGW_ZG2R_DB");
  const bool db = edb && atoi(edb) == 1;
  const char *edbt = getenv("SGW_ZG2R_DBT");                    // 1: two output tiles (A/B)
  const bool dbt = edbt && atoi(edbt) == 1 && !db;
  const size_t smem = g2r ? sizeof(cplx) * ((size_t)NZ * ZB * (dbt ? 2 : 1) + (size_t)a.maxlen * (db ? 2 : 1) + NZ) + sizeof(short) * ZB * RZ2 * ((RZ1 + 7) & ~7) + 32
                          : sizeof(cplx) * ((size_t)NZ * ZB + NZ) + (sizeof(double) + sizeof(short)) * (size_t)a.maxlen + 32;
  auto go = [&](auto kern) -> int {
    SGW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int vg = std::max(1, std::min(a.nvec, (per_sm * ctx->sm_count + ncb - 1) / ncb));
    dim3 grid((unsigned)ncb, (unsigned)vg);
    kern<<<grid, NT, smem, ctx->stream>>>(tm, a);
    return SGW_OK;
  };
  if (g2r) return db ? go(k_zpass_g2r_tma<RZ1, RZ2, NT, true>) : (dbt ? go(k_zpass_g2r_tma<RZ1, RZ2, NT, false, true>) : go(k_zpass_g2r_tma<RZ1, RZ2, NT, false>));
  if (a.maxlen <= 5 * NT) return go(k_zpass_r2g_tma<RZ1, RZ2, NT, 5>);
  return go(k_zpass_r2g_tma<RZ1, RZ2, NT, (ZB * NZ + NT - 1) / NT>);
}
static int zpass_tma(sgw_ctx *ctx, bool g2r, const GridDev &g, const Sphere &s, int nvec, const cplx *T, ZTArgs a, bool *done) {
  *done = false;
  if (!zpass_tma_ok(g, s)) return SGW_OK;
  const int var = zpass_variant();               // 1: both passes, 2: G -> r only, 3: r -> G only
  if ((var == 2 && !g2r) || (var == 3 && g2r)) return SGW_OK;
  CUtensorMap tm;
  if (!make_tmap_T(&tm, T, s.ncol, g.nz, nvec)) return SGW_OK;
  a.twz = g.twz; a.col_ptr = s.d_col_ptr; a.ztab = s.d_ztab; a.ncol = s.ncol; a.nvec = nvec; a.nz = g.nz; a.maxlen = s.zmaxlen;
  if (g.rz1 == 8) SGW_CHECK((launch_zpass_tma<8, 9>(ctx, g2r, tm, a)));
  else SGW_CHECK((launch_zpass_tma<5, 9>(ctx, g2r, tm, a)));
  *done = true;
  return SGW_OK;
}


// index table of the persistent plane kernels (k_plane_vloc, k_plane_rho_v2): for x column l (position in `xs`), sub-index j2 and
// butterfly leg k of the first y stage, the position of (xs[l], y = j2 + ry2 k) in a T row, or -1 outside the sphere
static int build_ytab(sgw_ctx *ctx, Sphere *sph, int nx, int ry1, int ry2, const std::vector<int> &xs) {
  if (sph->d_ytab) { dev_free(sph->d_ytab); sph->d_ytab = nullptr; }
  sph->ytab_ry1 = sph->ytab_ry2 = 0;
  if (ry2 <= 1 || sph->h_col_x.size() >= 32768) return SGW_OK;
  const int nxs = (int)xs.size();
  std::vector<int> xpos(nx, -1);
  for (int l = 0; l < nxs; ++l) xpos[xs[l]] = l;
  const int ry1p = (ry1 + 7) & ~7;                                // whole 16-byte loads per task
  std::vector<short> ytab((size_t)nxs * ry2 * ry1p, (short)-1);
  for (size_t c = 0; c < sph->h_col_x.size(); ++c) {
    const int l = xpos[sph->h_col_x[c]], y = sph->h_col_y[c];
    const int j2 = y % ry2, k = y / ry2;
    ytab[((size_t)j2 * nxs + l) * ry1p + k] = (short)c;
  }
  SGW_CHECK(upload(ctx, &sph->d_ytab, ytab.data(), ytab.size()));
  sph->ytab_ry1 = ry1; sph->ytab_ry2 = ry2;
  return SGW_OK;
}

// index table of the TMA z passes: offset of entry (column c, z) from the column's first entry, or -1 (upload; host arrays of `s`)
static int build_ztab(sgw_ctx *ctx, Sphere *s, int nz, int rz1, int rz2) {
  if (s->d_ztab) { dev_free(s->d_ztab); s->d_ztab = nullptr; }
  s->ztab_rz1 = s->ztab_rz2 = 0;
  if (rz2 <= 1 || nz != rz1 * rz2 || nz > 256) return SGW_OK;
  const int rz1p = (rz1 + 7) & ~7;
  std::vector<short> tab((size_t)s->ncol * rz2 * rz1p, (short)-1);
  int maxlen = 0;
  for (int c = 0; c < s->ncol; ++c)
    for (int p = s->h_col_ptr[c]; p < s->h_col_ptr[c + 1]; ++p) {
      const int z = s->h_zof[p];
      tab[((size_t)c * rz2 + z % rz2) * rz1p + z / rz2] = (short)(p - s->h_col_ptr[c]);
    }
  for (int c0 = 0; c0 < s->ncol; c0 += ZB) maxlen = std::max(maxlen, s->h_col_ptr[std::min(s->ncol, c0 + ZB)] - s->h_col_ptr[c0]);
  s->zmaxlen = (maxlen + 7) & ~7;
  SGW_CHECK(upload(ctx, &s->d_ztab, tab.data(), tab.size()));
  s->ztab_rz1 = rz1; s->ztab_rz2 = rz2;
  return SGW_OK;
}

int fft_zpass_g2r(sgw_ctx *ctx, const Sphere &s, int nvec, const cplx *in, long ld, cplx *T, const int *active,
                  const FftGrid *gr) {
  if (nvec <= 0) return SGW_OK;
  const GridDev g = grid_dev(ctx, gr);
  int zcb = zpass_zcb();
  if (zpass_smem(g, zcb) > ctx->smem_optin) zcb = 8;           // long lines (nz > ~880): fewer columns per CTA
  const size_t smem = zpass_smem(g, zcb);
  dim3 grid((s.ncol + zcb - 1) / zcb, nvec);
  ProfScope prof(ctx, ctx->prof_z_class);
  {
    ZTArgs a = {};
    a.vec = in; a.ld = ld; a.active = active;
    bool done = false;
    SGW_CHECK(zpass_tma(ctx, true, g, s, nvec, T, a, &done));
    if (done) { SGW_LAUNCH_CHECK(); return SGW_OK; }
  }
#define SGW_Z(Z)                                                                                       \
  do {                                                                                                 \
    SGW_CHECK(set_smem(ctx, k_zpass_g2r<Z>, smem));                                                    \
    k_zpass_g2r<Z><<<grid, zpass_threads(ctx), smem, ctx->stream>>>(g, s.dev(), in, ld, T, active);    \
  } while (0)
  if (zcb == 8) SGW_Z(8); else if (zcb == 32) SGW_Z(32); else SGW_Z(16);
#undef SGW_Z
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

int fft_zpass_r2g(sgw_ctx *ctx, const Sphere &s, int nvec, const cplx *T, cplx *out, long ld, const ZEpilogue &epi,
                  const int *active, const FftGrid *gr) {
  if (nvec <= 0) return SGW_OK;
  const GridDev g = grid_dev(ctx, gr);
  int zcb = zpass_zcb();
  if (zpass_smem(g, zcb) > ctx->smem_optin) zcb = 8;
  const size_t smem = zpass_smem(g, zcb);
  dim3 grid((s.ncol + zcb - 1) / zcb, nvec);
  const double scale = 1.0 / ((double)g.nx * g.ny * g.nz);
  ProfScope prof(ctx, ctx->prof_z_class);
  {
    ZTArgs a = {};
    a.vec = epi.psi; a.out = out; a.ld = ld; a.active = active;
    a.mode = epi.mode; a.keep_out = epi.keep_out; a.g2kin = epi.g2kin; a.sigma = epi.sigma; a.sigma_stride = epi.sigma_stride; a.scale = scale;
    bool done = false;
    SGW_CHECK(zpass_tma(ctx, false, g, s, nvec, T, a, &done));
    if (done) { SGW_LAUNCH_CHECK(); return SGW_OK; }
  }
#define SGW_Z(Z)                                                                                                  \
  do {                                                                                                            \
    SGW_CHECK(set_smem(ctx, k_zpass_r2g<Z>, smem));                                                               \
    k_zpass_r2g<Z><<<grid, zpass_threads(ctx), smem, ctx->stream>>>(g, s.dev(), T, out, ld, epi, scale, active);  \
  } while (0)
  static int ct = -1;                                           // SGW_ZR2G_CT=0: runtime radix dispatch also for 72 = 8 x 9 (A/B)
  if (ct < 0) { const char *e = getenv("SGW_ZR2G_CT"); ct = e ? atoi(e) : 1; }
  if (ct && zcb == 16 && g.rz1 == 8 && g.rz2 == 9) {
    SGW_CHECK(set_smem(ctx, k_zpass_r2g<16, 8, 9>, smem));
    k_zpass_r2g<16, 8, 9><<<grid, zpass_threads(ctx), smem, ctx->stream>>>(g, s.dev(), T, out, ld, epi, scale, active);
  } else if (zcb == 8) SGW_Z(8); else if (zcb == 32) SGW_Z(32); else SGW_Z(16);
#undef SGW_Z
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}


// ---- launch of the persistent VLOC kernel; returns SGW_OK and sets *done when an instantiation matches the grid
static int plane_vloc_variant() {
  // SGW_PLANE: 0 = generic k_plane, 1 = persistent + zero fill, 2 = persistent + x masks (default); read per call so
  // that the parity test can compare the variants inside one process
  const char *e = getenv("SGW_PLANE");
  return e ? atoi(e) : 2;
}
template <int RX1, int RX2, int RY1, int RY2, int NT>
static int launch_plane_vloc_nt(sgw_ctx *ctx, const GridDev &g, const Sphere &s, int nvec, const cplx *Tin, cplx *Tout, const int *active,
                                bool *done) {
  constexpr int NX = RX1 * RX2, NY = RY1 * RY2, PITCH = NX | 1;
  const int ncol_pad = (s.ncol + 7) & ~7, ntab = s.nxs * RY2 * ((RY1 + 7) & ~7);
  const size_t smem = sizeof(cplx) * ((size_t)NY * PITCH + ncol_pad + NX + NY) + sizeof(short) * ntab + sizeof(int) * 2 * NX +
                      sizeof(unsigned) * ((RX2 + 1) & ~1) + 16;
  if (smem > ctx->smem_optin) return SGW_OK;                    // not done: the generic kernel reports the limit
  PlaneVArgs a;
  a.twx = g.twx; a.twy = g.twy; a.Tin = Tin; a.Tout = Tout; a.vperm = ctx->d_vperm_t; a.active = active;
  a.xs = s.d_xs; a.ytab = s.d_ytab; a.nz = g.nz; a.nvec = nvec; a.ncol = s.ncol; a.nxs = s.nxs;
  { const char *e = getenv("SGW_YPAD"); a.ypad = e ? atoi(e) : 0; }
  const long nitems = (long)nvec * g.nz;
  auto go = [&](auto kern) -> int {
    SGW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const long ctas = std::min<long>(nitems, (long)per_sm * ctx->sm_count);
    kern<<<(unsigned)ctas, NT, smem, ctx->stream>>>(a);
    return SGW_OK;
  };
  if (plane_vloc_variant() == 1) SGW_CHECK(go(k_plane_vloc<RX1, RX2, RY1, RY2, NT, 1, false>));
  else SGW_CHECK(go(k_plane_vloc<RX1, RX2, RY1, RY2, NT, 1, true>));
  *done = true;
  return SGW_OK;
}
template <int RX1, int RX2, int RY1, int RY2>
static int launch_plane_vloc(sgw_ctx *ctx, const GridDev &g, const Sphere &s, int nvec, const cplx *Tin, cplx *Tout, const int *active,
                             bool *done) {
  // SGW_PLANE_NT: threads per CTA (192 | 224 | 256 | 288 | 320 | 352 | 384).  Measured at Si64 (fft_plane class per step, two CTAs per SM in
  // every case): 192 -> 100.2 ms, 224 -> 101.7 ms, 256 -> 101.6 ms, 320 -> 102.1 ms, 352 -> 104.7 ms, 384 -> 106.2 ms: fewer warps at
  // the block-wide barriers win, flat below 256
  const char *e = getenv("SGW_PLANE_NT");
  const int nt = e ? atoi(e) : 256;
  if (nt == 192) return launch_plane_vloc_nt<RX1, RX2, RY1, RY2, 192>(ctx, g, s, nvec, Tin, Tout, active, done);
  if (nt == 224) return launch_plane_vloc_nt<RX1, RX2, RY1, RY2, 224>(ctx, g, s, nvec, Tin, Tout, active, done);
  if (nt == 256) return launch_plane_vloc_nt<RX1, RX2, RY1, RY2, 256>(ctx, g, s, nvec, Tin, Tout, active, done);
  if (nt == 288) return launch_plane_vloc_nt<RX1, RX2, RY1, RY2, 288>(ctx, g, s, nvec, Tin, Tout, active, done);
  if (nt == 320) return launch_plane_vloc_nt<RX1, RX2, RY1, RY2, 320>(ctx, g, s, nvec, Tin, Tout, active, done);
  if (nt == 352) return launch_plane_vloc_nt<RX1, RX2, RY1, RY2, 352>(ctx, g, s, nvec, Tin, Tout, active, done);
  return launch_plane_vloc_nt<RX1, RX2, RY1, RY2, 384>(ctx, g, s, nvec, Tin, Tout, active, done);
}

// SGW_PLANE_GMEM=1 forces the global-memory plane path (testing it on boxes that would fit in shared memory)
static bool plane_force_gmem() {
  const char *e = getenv("SGW_PLANE_GMEM");
  return e && atoi(e) == 1;
}

static int try_plane_vloc(sgw_ctx *ctx, const GridDev &g, const Sphere &s, int nvec, const cplx *Tin, cplx *Tout, const int *active,
                          bool *done) {
  *done = false;
  if (plane_vloc_variant() == 0 || !s.d_ytab || s.ytab_ry1 != g.ry1 || s.ytab_ry2 != g.ry2) return SGW_OK;
#define SGW_PV(a, b, c, d) \
  if (g.rx1 == a && g.rx2 == b && g.ry1 == c && g.ry2 == d) return launch_plane_vloc<a, b, c, d>(ctx, g, s, nvec, Tin, Tout, active, done);
  SGW_PV(8, 9, 8, 9)        // 72 x 72 (Si64); other grids run the generic k_plane
#undef SGW_PV
  return SGW_OK;
}

int fft_plane(sgw_ctx *ctx, PlaneMode mode, const Sphere *sin, const Sphere *sout, int nvec, const cplx *Tin, cplx *Tout,
              const cplx *field, int vec_per_field, cplx *R, const int *active, int in_mod, const FftGrid *gr) {
  if (nvec <= 0) return SGW_OK;
  GridDev g = grid_dev(ctx, gr);
  const int P = planes_per_cta(g);
  const size_t smem = plane_smem(g, P);
  dim3 grid((g.nz + P - 1) / P, nvec);
  SphereDev si = sin ? sin->dev() : SphereDev(), so = sout ? sout->dev() : SphereDev();
  if (vec_per_field < 1) vec_per_field = 1;
  ProfScope prof(ctx, mode == PLANE_VLOC ? PC_FFT_PLANE : PC_OTHER);
  if (plane_smem(g, 1) > ctx->smem_optin || plane_force_gmem()) {
    // plane larger than an SM's shared memory: per-CTA planes in a global scratch of bounded size, vectors in chunks
    const size_t psz = (size_t)g.ny * g.pitchx;
    const size_t sm_tab = (size_t)(g.nx + g.ny) * sizeof(cplx) + 2 * (size_t)g.nx * sizeof(int);
    const size_t budget = (size_t)1 << 30;
    const int vchunk = (int)std::max<size_t>(1, std::min<size_t>({(size_t)nvec, (size_t)65535, budget / (psz * g.nz * sizeof(cplx))}));
    cplx *scratch = nullptr;
    SGW_CHECK(ws(ctx, "fft_plane_scratch", psz * g.nz * vchunk, &scratch));
    for (int v0 = 0; v0 < nvec; v0 += vchunk) {
      dim3 gg(g.nz, std::min(vchunk, nvec - v0));
#define SGW_PLANE_G(M)                                                                                                  \
      do {                                                                                                              \
        SGW_CHECK(set_smem(ctx, k_plane<M, 256, true, true>, sm_tab));                                                  \
        k_plane<M, 256, true, true><<<gg, 256, sm_tab, ctx->stream>>>(g, si, so, Tin, Tout, ctx->d_vperm, field, vec_per_field, R, in_mod, \
                                                                        active, 1, scratch, v0);                         \
      } while (0)
      switch (mode) {
        case PLANE_VLOC: SGW_PLANE_G(PLANE_VLOC); break;
        case PLANE_FIELD: SGW_PLANE_G(PLANE_FIELD); break;
        case PLANE_TO_R: SGW_PLANE_G(PLANE_TO_R); break;
        case PLANE_FROM_R: SGW_PLANE_G(PLANE_FROM_R); break;
      }
#undef SGW_PLANE_G
      SGW_LAUNCH_CHECK();
    }
    return SGW_OK;
  }
  if (mode == PLANE_VLOC && sin == sout && sin && P == 1 && !gr) {
    bool done = false;
    SGW_CHECK(try_plane_vloc(ctx, g, *sin, nvec, Tin, Tout, active, &done));
    if (done) {
      SGW_LAUNCH_CHECK();
      return SGW_OK;
    }
  }
#define SGW_PLANE_LAUNCH(M, NT)                                                                                          \
  do {                                                                                                                  \
    if (P == 1) {                                                                                                       \
      SGW_CHECK(set_smem(ctx, k_plane<M, NT, true>, smem));                                                             \
      k_plane<M, NT, true><<<grid, NT, smem, ctx->stream>>>(g, si, so, Tin, Tout, ctx->d_vperm, field, vec_per_field, R, in_mod, active, 1); \
    } else {                                                                                                            \
      SGW_CHECK(set_smem(ctx, k_plane<M, NT, false>, smem));                                                            \
      k_plane<M, NT, false><<<grid, NT, smem, ctx->stream>>>(g, si, so, Tin, Tout, ctx->d_vperm, field, vec_per_field, R, in_mod, active, P); \
    }                                                                                                                   \
  } while (0)
  const int nt = plane_threads();
  switch (mode) {
    case PLANE_VLOC:
      if (nt == 512) SGW_PLANE_LAUNCH(PLANE_VLOC, 512);
      else if (nt == 384) SGW_PLANE_LAUNCH(PLANE_VLOC, 384);
      else SGW_PLANE_LAUNCH(PLANE_VLOC, 256);
      break;
    case PLANE_FIELD: SGW_PLANE_LAUNCH(PLANE_FIELD, 256); break;
    case PLANE_TO_R: SGW_PLANE_LAUNCH(PLANE_TO_R, 256); break;
    case PLANE_FROM_R: SGW_PLANE_LAUNCH(PLANE_FROM_R, 256); break;
  }
#undef SGW_PLANE_LAUNCH
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

int fft_transpose_planes(sgw_ctx *ctx, const FftGrid *gr, long nplanes, const cplx *in, cplx *out) {
  const GridDev g = grid_dev(ctx, gr);
  for (long p0 = 0; p0 < nplanes; p0 += 65535) {
    const unsigned np = (unsigned)std::min<long>(65535, nplanes - p0);
    dim3 grid((g.nx + 15) / 16, (g.ny + 15) / 16, np);
    k_transpose_planes<<<grid, 256, 0, ctx->stream>>>(g.nx, g.ny, in + p0 * g.nx * g.ny, out + p0 * g.nx * g.ny);
    SGW_LAUNCH_CHECK();
  }
  return SGW_OK;
}

template <int RX1, int RX2, int RY1, int RY2>
static int launch_plane_rho_v2(sgw_ctx *ctx, const GridDev &g, const Sphere &sin, const Sphere &sout, int npf, int nocc, const cplx *Tin,
                               const cplx *psir_t, double wgt, cplx *Tout, int accumulate) {
  constexpr int NX = RX1 * RX2, NY = RY1 * RY2, PITCH = NX | 1, NT = 256;
  const int ncol_pad = (sin.ncol + 7) & ~7, ntab = sin.nxs * RY2 * ((RY1 + 7) & ~7);
  const size_t smem = sizeof(cplx) * ((size_t)NY * PITCH + ncol_pad + NX + NY) + sizeof(short) * ntab + sizeof(int) * 2 * NX +
                      sizeof(unsigned) * ((RX2 + 1) & ~1) + 16;
  if (smem > ctx->smem_optin) return 1;          // not done
  PlaneRhoArgs a;
  a.twx = g.twx; a.twy = g.twy; a.Tin = Tin; a.psir_t = psir_t; a.Tout = Tout; a.xs_in = sin.d_xs; a.xs_out = sout.d_xs;
  a.ytab_in = sin.d_ytab; a.col_off_out = sout.d_col_off; a.nz = g.nz; a.nocc = nocc; a.ncol_in = sin.ncol; a.nxs_in = sin.nxs;
  a.ncol_out = sout.ncol; a.nxs_out = sout.nxs; a.accumulate = accumulate; a.wgt = wgt;
  dim3 grid(npf, g.nz);
  // SGW_RHO_MINB: resident CTAs per SM the register allocation must allow.  Measured (Si64 step): 3 (80 registers, spills)
  // 25.4 ms, 2 (no spills) 20.5 ms; the generic k_plane_rho takes 40 ms
  const char *e = getenv("SGW_RHO_MINB");
  if (!(e && atoi(e) == 3)) {
    auto kern = k_plane_rho_v2<RX1, RX2, RY1, RY2, NT, 2>;
    SGW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, ctx->stream>>>(a);
  } else {
    auto kern = k_plane_rho_v2<RX1, RX2, RY1, RY2, NT, 3>;
    SGW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, NT, smem, ctx->stream>>>(a);
  }
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

int fft_plane_rho(sgw_ctx *ctx, const Sphere &sin, const Sphere &sout, int npf, int nocc, const cplx *Tin, const cplx *psir,
                  double wgt, cplx *Tout, int accumulate, const FftGrid *gr, const cplx *psir_t) {
  if (npf <= 0) return SGW_OK;
  const GridDev g = grid_dev(ctx, gr);
  const size_t smem = plane_smem(g, 2);
  dim3 grid(npf, g.nz);
  ProfScope prof(ctx, PC_RHO_PLANE);
  if (smem > ctx->smem_optin || plane_force_gmem()) {           // see fft_plane: planes in a bounded global scratch
    const size_t psz = 2 * (size_t)g.ny * g.pitchx;
    const size_t sm_tab = (size_t)(g.nx + g.ny) * sizeof(cplx) + 2 * (size_t)g.nx * sizeof(int);
    const size_t budget = (size_t)1 << 30;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)npf, budget / (psz * g.nz * sizeof(cplx))));
    cplx *scratch = nullptr;
    SGW_CHECK(ws(ctx, "fft_rho_scratch", psz * g.nz * chunk, &scratch));
    for (int p0 = 0; p0 < npf; p0 += chunk) {
      dim3 gg(std::min(chunk, npf - p0), g.nz);
      SGW_CHECK(set_smem(ctx, k_plane_rho<256, true>, sm_tab));
      k_plane_rho<256, true><<<gg, 256, sm_tab, ctx->stream>>>(g, sin.dev(), sout.dev(), nocc, Tin, psir, wgt, Tout, accumulate, scratch, p0);
      SGW_LAUNCH_CHECK();
    }
    return SGW_OK;
  }
  {
    const char *e = getenv("SGW_RHO_V2");          // 0: generic k_plane_rho (A/B testing)
    const bool want = !(e && atoi(e) == 0);
    if (want && psir_t && sin.d_ytab && sin.ytab_ry1 == g.ry1 && sin.ytab_ry2 == g.ry2 && g.rx1 == 5 && g.rx2 == 9 && g.ry1 == 5 &&
        g.ry2 == 9 && npf <= 2147483647 && g.nz <= 65535) {
      const int rc = launch_plane_rho_v2<5, 9, 5, 9>(ctx, g, sin, sout, npf, nocc, Tin, psir_t, wgt, Tout, accumulate);
      if (rc <= 0) return rc;
    }
  }
  // small planes (the reduced Delta-rho box): 256 threads fill the butterfly stages better and two CTAs fit on an SM
  static int rho_nt = -1;                                        // SGW_RHO_NT: tuning knob (128 | 192 | 256)
  if (rho_nt < 0) { const char *e = getenv("SGW_RHO_NT"); rho_nt = e ? atoi(e) : 0; }
  // measured on B200 (Si64, reduced 45^3 box): 256 threads / 2 CTAs per SM -> 48 ms per step, 192 -> 50 ms, 128 threads / 3 CTAs -> 40 ms
  if (g.nx * g.ny <= 3072 && (rho_nt == 128 || rho_nt == 0)) {
    SGW_CHECK(set_smem(ctx, k_plane_rho<128>, smem));
    k_plane_rho<128><<<grid, 128, smem, ctx->stream>>>(g, sin.dev(), sout.dev(), nocc, Tin, psir, wgt, Tout, accumulate);
  } else if (g.nx * g.ny <= 3072 && rho_nt == 192) {
    SGW_CHECK(set_smem(ctx, k_plane_rho<192>, smem));
    k_plane_rho<192><<<grid, 192, smem, ctx->stream>>>(g, sin.dev(), sout.dev(), nocc, Tin, psir, wgt, Tout, accumulate);
  } else if (g.nx * g.ny <= 3072) {   // SGW_RHO_NT=256
    SGW_CHECK(set_smem(ctx, k_plane_rho<256>, smem));
    k_plane_rho<256><<<grid, 256, smem, ctx->stream>>>(g, sin.dev(), sout.dev(), nocc, Tin, psir, wgt, Tout, accumulate);
  } else {
    SGW_CHECK(set_smem(ctx, k_plane_rho<512>, smem));
    k_plane_rho<512><<<grid, 512, smem, ctx->stream>>>(g, sin.dev(), sout.dev(), nocc, Tin, psir, wgt, Tout, accumulate);
  }
  SGW_LAUNCH_CHECK();
  return SGW_OK;
}

// Column structure of a sphere: entries sorted by (y, x, z); perm maps internal -> caller order.
int build_sphere(sgw_ctx *ctx, int npw, const int32_t *nl, Sphere *sph) {
  free_sphere(sph);
  const int nx = ctx->nr1, ny = ctx->nr2, nz = ctx->nr3;
  std::vector<int> order(npw);
  std::vector<long> key(npw);
  for (int i = 0; i < npw; ++i) {
    const long idx = (long)nl[i] - 1;
    if (idx < 0 || idx >= (long)nx * ny * nz) {
      ctx->err = "nl index outside the FFT box";
      return SGW_E_ARG;
    }
    const long x = idx % nx, y = (idx / nx) % ny, z = idx / ((long)nx * ny);
    key[i] = (y * nx + x) * nz + z;
    order[i] = i;
  }
  std::sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  for (int i = 1; i < npw; ++i)
    if (key[order[i]] == key[order[i - 1]]) {
      ctx->err = "duplicate nl index in plane-wave list";
      return SGW_E_ARG;
    }
  std::vector<int> col_x, col_y, col_ptr, colof(npw), zof(npw);
  std::vector<char> xused(nx, 0);
  long last = -1;
  for (int p = 0; p < npw; ++p) {
    const long k = key[order[p]];
    const long colkey = k / nz;
    if (colkey != last) {
      col_ptr.push_back(p);
      col_x.push_back((int)(colkey % nx));
      col_y.push_back((int)(colkey / nx));
      xused[colkey % nx] = 1;
      last = colkey;
    }
    colof[p] = (int)col_x.size() - 1;
    zof[p] = (int)(k % nz);
  }
  col_ptr.push_back(npw);
  std::vector<int> xs;
  for (int x = 0; x < nx; ++x)
    if (xused[x]) xs.push_back(x);
  {
    // start the list after the largest (cyclic) gap: a sphere occupies x = -r..r, i.e. [nx-r..nx-1, 0..r]; in that order
    // consecutive entries are consecutive mod nx, so the lanes of a warp that walk the list touch consecutive 16-byte
    // bank groups across the wrap as well (nx = 0 mod 8 on the usual grids)
    size_t best = 0;
    int gap = -1;
    for (size_t i = 0; i < xs.size(); ++i) {
      const int prev = xs[(i + xs.size() - 1) % xs.size()];
      const int d = ((xs[i] - prev) % nx + nx) % nx;
      if (d > gap) { gap = d; best = i; }
    }
    { const char *e = getenv("SGW_XROT"); if (!e || atoi(e)) std::rotate(xs.begin(), xs.begin() + best, xs.end()); }
  }
  sph->npw = npw;
  sph->ncol = (int)col_x.size();
  sph->nxs = (int)xs.size();
  sph->perm = order;
  sph->h_col_x = col_x; sph->h_col_y = col_y; sph->h_col_ptr = col_ptr; sph->h_zof = zof;
  SGW_CHECK(upload(ctx, &sph->d_col_x, col_x.data(), col_x.size()));
  SGW_CHECK(upload(ctx, &sph->d_col_y, col_y.data(), col_y.size()));
  {
    std::vector<int> col_off(col_x.size());
    for (size_t c = 0; c < col_x.size(); ++c) col_off[c] = col_y[c] * (nx | 1) + col_x[c];
    SGW_CHECK(upload(ctx, &sph->d_col_off, col_off.data(), col_off.size()));
  }
  SGW_CHECK(upload(ctx, &sph->d_col_ptr, col_ptr.data(), col_ptr.size()));
  SGW_CHECK(upload(ctx, &sph->d_colof, colof.data(), colof.size()));
  SGW_CHECK(upload(ctx, &sph->d_zof, zof.data(), zof.size()));
  SGW_CHECK(upload(ctx, &sph->d_xs, xs.data(), xs.size()));
  SGW_CHECK(upload(ctx, &sph->d_perm, order.data(), order.size()));
  SGW_CHECK(build_ytab(ctx, sph, nx, ctx->py.r1, ctx->py.r2, xs));
  SGW_CHECK(build_ztab(ctx, sph, nz, ctx->pz.r1, ctx->pz.r2));
  return SGW_OK;
}

// Re-express a sphere in another box: every entry keeps its Miller indices (box coordinate c of an n-point axis
// means m = c for c <= (n-1)/2, else c - n), its position in the vector and its column; only the box coordinates
// change, so coefficient vectors are shared between the two grids.  Fails if the box is too small for the sphere.
int remap_sphere(sgw_ctx *ctx, const Sphere &fine, const FftGrid &gr, Sphere *out) {
  free_sphere(out);
  const int nf[3] = {ctx->nr1, ctx->nr2, ctx->nr3}, nc[3] = {gr.n1, gr.n2, gr.n3};
  auto conv = [&](int c, int d, int *res) {
    const int m = c <= (nf[d] - 1) / 2 ? c : c - nf[d];
    if (2 * std::abs(m) >= nc[d]) return false;
    *res = ((m % nc[d]) + nc[d]) % nc[d];
    return true;
  };
  const int ncol = fine.ncol, npw = fine.npw;
  std::vector<int> col_x(ncol), col_y(ncol), col_off(ncol), zof(npw), colof(npw);
  std::vector<char> xused(nc[0], 0);
  bool ok = true;
  for (int c = 0; c < ncol; ++c) {
    ok = ok && conv(fine.h_col_x[c], 0, &col_x[c]) && conv(fine.h_col_y[c], 1, &col_y[c]);
    if (!ok) break;
    col_off[c] = col_y[c] * (nc[0] | 1) + col_x[c];
    xused[col_x[c]] = 1;
    for (int p = fine.h_col_ptr[c]; p < fine.h_col_ptr[c + 1]; ++p) colof[p] = c;
  }
  for (int p = 0; ok && p < npw; ++p) ok = conv(fine.h_zof[p], 2, &zof[p]);
  if (!ok) {
    ctx->err = "remap_sphere: target FFT box too small for the sphere";
    return SGW_E_ARG;
  }
  std::vector<int> xs;
  for (int x = 0; x < nc[0]; ++x)
    if (xused[x]) xs.push_back(x);
  out->npw = npw; out->ncol = ncol; out->nxs = (int)xs.size();
  out->perm = fine.perm;
  out->h_col_x = col_x; out->h_col_y = col_y; out->h_col_ptr = fine.h_col_ptr; out->h_zof = zof;
  SGW_CHECK(upload(ctx, &out->d_col_x, col_x.data(), col_x.size()));
  SGW_CHECK(upload(ctx, &out->d_col_y, col_y.data(), col_y.size()));
  SGW_CHECK(upload(ctx, &out->d_col_off, col_off.data(), col_off.size()));
  SGW_CHECK(upload(ctx, &out->d_col_ptr, fine.h_col_ptr.data(), fine.h_col_ptr.size()));
  SGW_CHECK(upload(ctx, &out->d_colof, colof.data(), colof.size()));
  SGW_CHECK(upload(ctx, &out->d_zof, zof.data(), zof.size()));
  SGW_CHECK(upload(ctx, &out->d_xs, xs.data(), xs.size()));
  SGW_CHECK(upload(ctx, &out->d_perm, fine.perm.data(), fine.perm.size()));
  SGW_CHECK(build_ztab(ctx, out, gr.n3, gr.pz.r1, gr.pz.r2));
  SGW_CHECK(build_ytab(ctx, out, nc[0], gr.py.r1, gr.py.r2, xs));
  return SGW_OK;
}

void free_sphere(Sphere *s) {
  int **ptrs[] = {&s->d_col_x, &s->d_col_y, &s->d_col_ptr, &s->d_colof, &s->d_zof, &s->d_xs, &s->d_perm, &s->d_col_off};
  for (auto p : ptrs) {
    if (*p) dev_free(*p);
    *p = nullptr;
  }
  if (s->d_ytab) dev_free(s->d_ytab);
  s->d_ytab = nullptr;
  if (s->d_ztab) dev_free(s->d_ztab);
  s->d_ztab = nullptr;
  s->ztab_rz1 = s->ztab_rz2 = 0;
  s->ytab_ry1 = s->ytab_ry2 = 0;
  s->npw = s->ncol = s->nxs = 0;
  s->perm.clear();
  s->h_col_x.clear(); s->h_col_y.clear(); s->h_col_ptr.clear(); s->h_zof.clear();
}

}  // namespace sgw
