// CPU lane-by-lane emulation of the blind-rotation kernel's data flow, built from the SAME
// __host__ __device__ phase functions the CUDA kernel uses (fhe_regex_b200/csrc/br_core.cuh).
// Test infrastructure: lets the index / twiddle / swizzle logic be checked against the oracle in
// the build container, which has no GPU.  Not part of the product.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fhe_regex_b200/csrc/br_core.cuh"

using namespace fb;

struct Regs { double xr[32], xi[32]; };

struct Sample {
  std::vector<uint64_t> acc;   // [2][2048]
  std::vector<double> plane;   // [2][1024]: the re and the im planes pass through it one after the other
  Regs regs[2][32];            // [warp][lane]
  Sample() : acc(2 * kN), plane(2 * kHalfN) {}
};

static c2 g_tab_f[kTabEntries * 32], g_tab_i[kTabEntries * 32];
static bool g_tabs = false;
static void tabs() { if (!g_tabs) { make_twiddle_tables(g_tab_f, g_tab_i); g_tabs = true; } }

// forward half shared by the key conversion and the CMUX: phases A2-A3, barrier, B1-B2
static void forward_passes(Sample& s) {
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      Regs& R = s.regs[w][lane];
      fft32_dif(R.xr, R.xi);
      fwd_twiddle_inplace(R.xr, R.xi, g_tab_f, lane);
    }
  // barrier; re plane
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_store_brev(s.regs[w][lane].xr, s.plane.data() + w * kHalfN, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_load(s.regs[w][lane].xr, s.plane.data() + (lane >> 4) * kHalfN, 16 * w + (lane & 15));
  // barrier; im plane
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_store_brev(s.regs[w][lane].xi, s.plane.data() + w * kHalfN, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_load(s.regs[w][lane].xi, s.plane.data() + (lane >> 4) * kHalfN, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) fft32_dif(s.regs[w][lane].xr, s.regs[w][lane].xi);
}

// inverse half: inverse pass 1, twiddle, split transpose, inverse pass 2 (leaves phase-C input in regs)
static void inverse_passes(Sample& s) {
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      Regs& R = s.regs[w][lane];
      fft32_dit_inv(R.xr, R.xi);
      inv_twiddle_inplace(R.xr, R.xi, g_tab_i, 16 * w + (lane & 15));
    }
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_store(s.regs[w][lane].xr, s.plane.data() + (lane >> 4) * kHalfN, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_load_brev(s.regs[w][lane].xr, s.plane.data() + w * kHalfN, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_store(s.regs[w][lane].xi, s.plane.data() + (lane >> 4) * kHalfN, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_load_brev(s.regs[w][lane].xi, s.plane.data() + w * kHalfN, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) fft32_dit_inv(s.regs[w][lane].xr, s.regs[w][lane].xi);
}

extern "C" void emu_bsk_to_fourier(const uint64_t* bsk, c2* fbsk) {
  tabs();
  Sample s;
  for (int t = 0; t < kLweN * 2; t++) {  // (i, row): two polys each
    for (int w = 0; w < 2; w++)
      for (int lane = 0; lane < 32; lane++)
        phaseA_load_torus(s.regs[w][lane].xr, s.regs[w][lane].xi, bsk + ((size_t)t * 2 + w) * kN, lane);
    forward_passes(s);
    for (int w = 0; w < 2; w++)
      for (int lane = 0; lane < 32; lane++) {
        Regs& R = s.regs[w][lane];
        const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
        for (int q = 0; q < 32; q++) {
          const int k = k1 + 32 * brev5(q);
          c2 v; v.x = R.xr[q]; v.y = R.xi[q];
          fbsk[((size_t)t * 2 + pp) * kHalfN + k] = v;
        }
      }
  }
}

static void cmux_step(Sample& s, const c2* fbsk, int i, uint32_t a) {
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++)
      phaseA_load(s.regs[w][lane].xr, s.regs[w][lane].xi, s.acc.data() + w * kN, a, lane);
  forward_passes(s);
  // MAC with the shuffle exchange between lane and lane^16
  for (int w = 0; w < 2; w++) {
    for (int q = 0; q < 32; q++) {
      double keep_r[32], keep_i[32], send_r[32], send_i[32];
      for (int lane = 0; lane < 32; lane++) {
        Regs& R = s.regs[w][lane];
        const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
        const int k = k1 + 32 * brev5(q);
        mac_point(R.xr[q], R.xi[q], fbsk[fbsk_index(i, pp, pp, k)], fbsk[fbsk_index(i, pp, 1 - pp, k)],
                  keep_r[lane], keep_i[lane], send_r[lane], send_i[lane]);
      }
      for (int lane = 0; lane < 32; lane++) {
        s.regs[w][lane].xr[q] = keep_r[lane] + send_r[lane ^ 16];
        s.regs[w][lane].xi[q] = keep_i[lane] + send_i[lane ^ 16];
      }
    }
  }
  inverse_passes(s);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++)
      phaseC_update(s.regs[w][lane].xr, s.regs[w][lane].xi, s.acc.data() + w * kN, lane);
}

// small[743], lut[2048] -> acc[2][2048]; max_steps < 0 means all 742
extern "C" void emu_blind_rotate(const c2* fbsk, const uint64_t* small, const uint64_t* lut, uint64_t* acc_out, int max_steps) {
  tabs();
  Sample s;
  const uint32_t bt = modswitch(small[kLweN]);
  for (int j = 0; j < kN; j++) {
    s.acc[j] = 0;
    s.acc[kN + j] = rot_read(lut, j, (4096u - bt) & 4095u);
  }
  const int steps = max_steps < 0 ? kLweN : max_steps;
  for (int i = 0; i < steps; i++) {
    if (small[i] == 0) continue;
    cmux_step(s, fbsk, i, modswitch(small[i]) & 4095u);
  }
  memcpy(acc_out, s.acc.data(), sizeof(uint64_t) * 2 * kN);
}

// negacyclic product check: out = round(a_int (*) b_torus) using forward/pointwise/inverse of the emulated passes
extern "C" void emu_negacyclic_mul(const int64_t* a_int, const uint64_t* b_torus, uint64_t* out) {
  tabs();
  Sample sa, sb;
  std::vector<uint64_t> zero(kN, 0);
  for (int lane = 0; lane < 32; lane++) {
    // warp 0 <- polynomial under test, warp 1 <- zero
    for (int r = 0; r < 32; r++) {
      const int j = 32 * r + lane;
      const double d0 = (double)a_int[j], d1 = (double)a_int[j + 1024];
      const double cr = fb_twist_cos(r), sr = fb_twist_sin(r);
      sa.regs[0][lane].xr[r] = d0 * cr - d1 * sr; sa.regs[0][lane].xi[r] = d0 * sr + d1 * cr;
      sa.regs[1][lane].xr[r] = 0; sa.regs[1][lane].xi[r] = 0;
    }
    phaseA_load_torus(sb.regs[0][lane].xr, sb.regs[0][lane].xi, b_torus, lane);
    phaseA_load_torus(sb.regs[1][lane].xr, sb.regs[1][lane].xi, zero.data(), lane);
  }
  forward_passes(sa);
  forward_passes(sb);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      Regs &A = sa.regs[w][lane], &B = sb.regs[w][lane];
      for (int q = 0; q < 32; q++) {
        double r = A.xr[q] * B.xr[q] - A.xi[q] * B.xi[q], im = A.xr[q] * B.xi[q] + A.xi[q] * B.xr[q];
        A.xr[q] = r; A.xi[q] = im;
      }
    }
  inverse_passes(sa);
  std::fill(sa.acc.begin(), sa.acc.end(), 0);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      Regs& R = sa.regs[w][lane];
      phaseC_update(R.xr, R.xi, sa.acc.data() + w * kN, lane);
    }
  memcpy(out, sa.acc.data(), sizeof(uint64_t) * kN);
}

// test hook: raw forward passes on already folded+twisted-by-register input z[32*r+lane] (poly 0 slot)
extern "C" void emu_forward_raw(const double* zr, const double* zi, double* outr, double* outi) {
  tabs();
  Sample s;
  for (int lane = 0; lane < 32; lane++)
    for (int r = 0; r < 32; r++) {
      s.regs[0][lane].xr[r] = zr[32 * r + lane]; s.regs[0][lane].xi[r] = zi[32 * r + lane];
      s.regs[1][lane].xr[r] = 0; s.regs[1][lane].xi[r] = 0;
    }
  forward_passes(s);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 16; lane++) {
      const int k1 = 16 * w + lane;
      for (int q = 0; q < 32; q++) { outr[k1 + 32 * brev5(q)] = s.regs[w][lane].xr[q]; outi[k1 + 32 * brev5(q)] = s.regs[w][lane].xi[q]; }
    }
}
