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
  std::vector<uint64_t> acc;   // [2][2048] (only used by the product check)
  std::vector<uint32_t> shadow; // [2][2048] 32-bit accumulator, shared copy (rotation reads of the decomposition)
  uint32_t master[2][32][64];  // [warp][lane][2r + half]: thread-private copy of the same words (tensor memory on the device)
  std::vector<double> plane;   // [2][kPlaneDoubles]: the re and the im planes pass through it one after the other
  Regs regs[2][32];            // [warp][lane]
  Sample() : acc(2 * kN), shadow(2 * kN), plane(2 * kPlaneDoubles) {}
};

static c2 g_tab_f[kTabEntries * 32], g_tab_i[kTabEntries * 32];
static bool g_tabs = false;
static void tabs() { if (!g_tabs) { make_twiddle_tables(g_tab_f, g_tab_i); g_tabs = true; } }

// forward half shared by the key conversion and the CMUX: phases A2-A3, barrier, B1-B2
static void forward_passes(Sample& s) {
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      Regs& R = s.regs[w][lane];
      fft32_fwd_twist(R.xr, R.xi);   // input: untwisted folded coefficients
      fwd_twiddle_inplace(R.xr, R.xi, g_tab_f, lane);
    }
  // barrier; re plane
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_store_brev(s.regs[w][lane].xr, s.plane.data() + w * kPlaneDoubles, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_load(s.regs[w][lane].xr, s.plane.data() + (lane >> 4) * kPlaneDoubles, 16 * w + (lane & 15));
  // barrier; im plane
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_store_brev(s.regs[w][lane].xi, s.plane.data() + w * kPlaneDoubles, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_load(s.regs[w][lane].xi, s.plane.data() + (lane >> 4) * kPlaneDoubles, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) fft32_fwd(s.regs[w][lane].xr, s.regs[w][lane].xi);
}

// inverse half: inverse pass 1, twiddle, split transpose, inverse pass 2 (leaves phase-C input in regs)
static void inverse_passes(Sample& s) {
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      Regs& R = s.regs[w][lane];
      fft32_inv(R.xr, R.xi);
      inv_twiddle_inplace(R.xr, R.xi, g_tab_i, 16 * w + (lane & 15));
    }
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_store(s.regs[w][lane].xr, s.plane.data() + (lane >> 4) * kPlaneDoubles, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_load_brev(s.regs[w][lane].xr, s.plane.data() + w * kPlaneDoubles, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_store(s.regs[w][lane].xi, s.plane.data() + (lane >> 4) * kPlaneDoubles, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_load_brev(s.regs[w][lane].xi, s.plane.data() + w * kPlaneDoubles, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) fft32_inv(s.regs[w][lane].xr, s.regs[w][lane].xi);
}

extern "C" void emu_bsk_to_fourier(const uint64_t* bsk, c2* fbsk) {
  tabs();
  Sample s;
  for (int t = 0; t < kLweN * 2; t++) {  // (i, row): two polys each
    for (int w = 0; w < 2; w++)
      for (int lane = 0; lane < 32; lane++)
        load_torus_poly(s.regs[w][lane].xr, s.regs[w][lane].xi, bsk + ((size_t)t * 2 + w) * kN, lane);
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
      phaseA_load32(s.regs[w][lane].xr, s.regs[w][lane].xi, s.shadow.data() + w * kN, a, lane);
  forward_passes(s);
  // MAC: every lane takes the other polynomial's spectrum value from lane ^ 16
  for (int w = 0; w < 2; w++) {
    for (int q = 0; q < 32; q++) {
      double pr[32], pi[32];
      for (int lane = 0; lane < 32; lane++) { pr[lane] = s.regs[w][lane ^ 16].xr[q]; pi[lane] = s.regs[w][lane ^ 16].xi[q]; }
      for (int lane = 0; lane < 32; lane++) {
        Regs& R = s.regs[w][lane];
        const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
        const int k = k1 + 32 * brev5(q);
        mac_point2(R.xr[q], R.xi[q], pr[lane], pi[lane], fbsk[fbsk_index(i, pp, pp, k)], fbsk[fbsk_index(i, 1 - pp, pp, k)]);
      }
    }
  }
  inverse_passes(s);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++)
      for (int r = 0; r < 32; r++) {
        uint32_t inc0, inc1;
        phaseC_increments32(s.regs[w][lane].xr, s.regs[w][lane].xi, r, inc0, inc1);
        uint32_t& a0 = s.master[w][lane][2 * r];
        uint32_t& a1 = s.master[w][lane][2 * r + 1];
        a0 += inc0;
        a1 += inc1;
        s.shadow[w * kN + 32 * r + lane] = a0;
        s.shadow[w * kN + 32 * r + lane + 1024] = a1;
      }
}

// ---- the fused CMUX body of br_fused.cu, piece by piece in the kernel's order ------------------------------
template <int B>
static void emu_mid_block(Sample& s, const c2* fbsk, int i) {
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      Regs& R = s.regs[w][lane];
      fft32_fwd_s3<B>(R.xr, R.xi);
      fft32_fwd_s45<2 * B>(R.xr, R.xi);
      fft32_fwd_s45<2 * B + 1>(R.xr, R.xi);
    }
  for (int w = 0; w < 2; w++)
    for (int t = 0; t < 8; t++) {
      const int q = 8 * B + t;
      double pr[32], pi[32];
      for (int lane = 0; lane < 32; lane++) { pr[lane] = s.regs[w][lane ^ 16].xr[q]; pi[lane] = s.regs[w][lane ^ 16].xi[q]; }
      for (int lane = 0; lane < 32; lane++) {
        Regs& R = s.regs[w][lane];
        const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
        const int k = k1 + 32 * brev5(q);
        mac_point2(R.xr[q], R.xi[q], pr[lane], pi[lane], fbsk[fbsk_index(i, pp, pp, k)], fbsk[fbsk_index(i, 1 - pp, pp, k)]);
      }
    }
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      Regs& R = s.regs[w][lane];
      fft32_inv_s12<2 * B>(R.xr, R.xi);
      fft32_inv_s12<2 * B + 1>(R.xr, R.xi);
      fft32_inv_s3<B>(R.xr, R.xi);
    }
}
template <int A>
static void emu_fin_pair(Sample& s, int w, int lane) {
  Regs& R = s.regs[w][lane];
  fft32_i2_fin<A>(R.xr, R.xi);
  uint32_t i0, i1;
  phaseC_slot<A>(R.xr, R.xi, i0, i1);
  s.master[w][lane][2 * A] += i0;
  s.master[w][lane][2 * A + 1] += i1;
  phaseC_slot<A + 16>(R.xr, R.xi, i0, i1);
  s.master[w][lane][2 * (A + 16)] += i0;
  s.master[w][lane][2 * (A + 16) + 1] += i1;
}
template <int... As>
static void emu_fin_all(Sample& s, int w, int lane, fb_iseq<As...>) { (emu_fin_pair<As>(s, w, lane), ...); }

static void cmux_step_fused(Sample& s, const c2* fbsk, int i, uint32_t a) {
  // the device passes the shared-memory block and the byte offset of the polynomial's copy (a multiple of 8 KiB)
  const unsigned char* sm = reinterpret_cast<const unsigned char*>(s.shadow.data());
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) phaseA_f1(s.regs[w][lane].xr, s.regs[w][lane].xi, sm, (uint32_t)w * 8192u, a, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) fwd_twiddle_inplace(s.regs[w][lane].xr, s.regs[w][lane].xi, g_tab_f, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_store_brev(s.regs[w][lane].xr, s.plane.data() + w * kPlaneDoubles, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_load(s.regs[w][lane].xr, s.plane.data() + (lane >> 4) * kPlaneDoubles, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_store_brev(s.regs[w][lane].xi, s.plane.data() + w * kPlaneDoubles, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_load(s.regs[w][lane].xi, s.plane.data() + (lane >> 4) * kPlaneDoubles, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) fft32_fwd_s12(s.regs[w][lane].xr, s.regs[w][lane].xi);
  emu_mid_block<0>(s, fbsk, i);
  emu_mid_block<1>(s, fbsk, i);
  emu_mid_block<2>(s, fbsk, i);
  emu_mid_block<3>(s, fbsk, i);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      Regs& R = s.regs[w][lane];
      fft32_inv_s45(R.xr, R.xi);
      inv_twiddle_inplace(R.xr, R.xi, g_tab_i, 16 * w + (lane & 15));
    }
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_store(s.regs[w][lane].xr, s.plane.data() + (lane >> 4) * kPlaneDoubles, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_load_brev(s.regs[w][lane].xr, s.plane.data() + w * kPlaneDoubles, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) row_store(s.regs[w][lane].xi, s.plane.data() + (lane >> 4) * kPlaneDoubles, 16 * w + (lane & 15));
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) col_load_brev(s.regs[w][lane].xi, s.plane.data() + w * kPlaneDoubles, lane);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++) {
      fft32_i2_head(s.regs[w][lane].xr, s.regs[w][lane].xi);
      emu_fin_all(s, w, lane, fb_make_iseq<16>{});
    }
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++)
      for (int r = 0; r < 32; r++) {
        s.shadow[w * kN + 32 * r + lane] = s.master[w][lane][2 * r];
        s.shadow[w * kN + 32 * r + lane + 1024] = s.master[w][lane][2 * r + 1];
      }
}

static bool g_fused = false;
extern "C" void emu_set_fused(int on) { g_fused = on != 0; }

// small[743], lut[2048] -> acc[2][2048]; max_steps < 0 means all 742
extern "C" void emu_blind_rotate(const c2* fbsk, const uint64_t* small, const uint64_t* lut, uint64_t* acc_out, int max_steps) {
  tabs();
  Sample& s = *new Sample();
  const uint32_t bt = modswitch(small[kLweN]);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++)
      for (int r = 0; r < 32; r++)
        for (int h = 0; h < 2; h++) {
          const int j = 32 * r + lane + 1024 * h;
          const uint32_t v = w == 0 ? 0u : (uint32_t)(rot_read(lut, j, (4096u - bt) & 4095u) >> 32);
          s.master[w][lane][2 * r + h] = v;
          s.shadow[w * kN + j] = v;
        }
  const int steps = max_steps < 0 ? kLweN : max_steps;
  for (int i = 0; i < steps; i++) {
    const uint32_t a = modswitch(small[i]) & 4095u;
    if (small[i] == 0 || a == 0) continue;
    if (g_fused) cmux_step_fused(s, fbsk, i, a);
    else cmux_step(s, fbsk, i, a);
  }
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++)
      for (int r = 0; r < 32; r++)
        for (int h = 0; h < 2; h++) acc_out[w * kN + 32 * r + lane + 1024 * h] = (uint64_t)s.master[w][lane][2 * r + h] << 32;
  delete &s;
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
      sa.regs[0][lane].xr[r] = (double)a_int[j]; sa.regs[0][lane].xi[r] = (double)a_int[j + 1024];
      sa.regs[1][lane].xr[r] = 0; sa.regs[1][lane].xi[r] = 0;
    }
    load_torus_poly(sb.regs[0][lane].xr, sb.regs[0][lane].xi, b_torus, lane);
    load_torus_poly(sb.regs[1][lane].xr, sb.regs[1][lane].xi, zero.data(), lane);
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
      for (int r = 0; r < 32; r++) {
        uint32_t inc0, inc1;
        phaseC_increments32(R.xr, R.xi, r, inc0, inc1);
        sa.acc[w * kN + 32 * r + lane] = (uint64_t)inc0 << 32;
        sa.acc[w * kN + 32 * r + lane + 1024] = (uint64_t)inc1 << 32;
      }
    }
  memcpy(out, sa.acc.data(), sizeof(uint64_t) * kN);
}


// ---- round 2, last session: the transposes through the other plane layouts ------------------------------------------
// One sample's worth of registers goes through the forward twiddle + transpose and the inverse twiddle + transpose in
//   layout 0: one contiguous plane for both components, one after the other (col_store_brev / row_load / row_store / col_load_brev)
//   layout 1: planes inside the 8 KiB accumulator copies + overflow blocks (col_store_brev_al / plane_al_row / col_load_brev_al)
//   layout 2: a plane per component with the twiddles fused into the stores (fwd_twiddle_col_store / inv_twiddle_row_store)
//   layout 3: layout 1 with the full twiddle tables (fb_full_twiddle / fwd_twiddle_full / inv_twiddle_full)
// returns the number of registers that differ from layout 0 bit for bit (0 expected), or -1 if a layout wrote outside its blocks
extern "C" int emu_plane_layouts_check(int layout, unsigned seed) {
  tabs();
  static Regs ref[2][32], got[2][32], in[2][32];
  srand(seed);
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++)
      for (int r = 0; r < 32; r++) {
        in[w][lane].xr[r] = (double)(rand() % 2000001 - 1000000) / 7.0;
        in[w][lane].xi[r] = (double)(rand() % 2000001 - 1000000) / 11.0;
      }
  auto run = [&](int L, Regs (&R)[2][32]) -> bool {
    memcpy(R, in, sizeof(in));
    // shared memory of one sample: two accumulator copies of 8 KiB, two overflow blocks, two imaginary planes, guard words between
    const size_t kGuard = 16;
    std::vector<double> shadow(2 * 1024 + kGuard, -7.0), ovf(2 * (kPlaneOvfBytes / 8) + kGuard, -7.0), implane(2 * kPlaneDoubles + kGuard, -7.0),
        plane(2 * kPlaneDoubles + kGuard, -7.0);
    std::vector<c2> full_f(1024), full_i(1024);
    for (int t = 0; t < 1024; t++) {
      full_f[t] = fb_full_twiddle(g_tab_f, t >> 5, t & 31);
      full_i[t] = fb_full_twiddle(g_tab_i, t >> 5, t & 31);
    }
    auto main_of = [&](int p) { return shadow.data() + p * 1024; };
    auto ovf_of = [&](int p) { return ovf.data() + p * (kPlaneOvfBytes / 8) + kPlaneOvfLead / 8; };
    auto pass = [&](bool inverse) {
      for (int comp = 0; comp < 2; comp++) {   // layouts 0, 1, 3: the components one after the other
        if (L == 2 && comp == 1) break;
        for (int w = 0; w < 2; w++)
          for (int lane = 0; lane < 32; lane++) {
            Regs& T = R[w][lane];
            const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
            double (&x)[32] = comp ? T.xi : T.xr;
            if (!inverse) {
              if (comp == 0 && L != 2) { if (L == 3) fwd_twiddle_full(T.xr, T.xi, full_f.data() + lane); else fwd_twiddle_inplace(T.xr, T.xi, g_tab_f, lane); }
              if (L == 0) col_store_brev(x, plane.data() + w * kPlaneDoubles, lane);
              else if (L == 2) fwd_twiddle_col_store(T.xr, T.xi, g_tab_f, lane, main_of(w) + lane, ovf_of(w) + lane, implane.data() + w * kPlaneDoubles + lane);
              else col_store_brev_al(x, main_of(w) + lane, ovf_of(w) + lane);
            } else {
              if (comp == 0 && L != 2) { if (L == 3) inv_twiddle_full(T.xr, T.xi, full_i.data() + k1); else inv_twiddle_inplace(T.xr, T.xi, g_tab_i, k1); }
              if (L == 0) row_store(x, plane.data() + pp * kPlaneDoubles, k1);
              else if (L == 2) inv_twiddle_row_store(T.xr, T.xi, g_tab_i, k1, plane_al_row(main_of(pp), ovf_of(pp), k1), implane.data() + pp * kPlaneDoubles + k1 * kPlaneRow);
              else row_store(x, plane_al_row(main_of(pp), ovf_of(pp), k1), 0);
            }
          }
        for (int cc = comp; cc < (L == 2 ? 2 : comp + 1); cc++)   // barrier; the readers
          for (int w = 0; w < 2; w++)
            for (int lane = 0; lane < 32; lane++) {
              Regs& T = R[w][lane];
              const int pp = lane >> 4, k1 = 16 * w + (lane & 15);
              double (&x)[32] = cc ? T.xi : T.xr;
              if (!inverse) {
                if (L == 0) row_load(x, plane.data() + pp * kPlaneDoubles, k1);
                else if (L == 2 && cc == 1) row_load(x, implane.data() + pp * kPlaneDoubles + k1 * kPlaneRow, 0);
                else row_load(x, plane_al_row(main_of(pp), ovf_of(pp), k1), 0);
              } else {
                if (L == 0) col_load_brev(x, plane.data() + w * kPlaneDoubles, lane);
                else if (L == 2 && cc == 1) col_load_brev(x, implane.data() + w * kPlaneDoubles + lane, 0);
                else col_load_brev_al(x, main_of(w) + lane, ovf_of(w) + lane);
              }
            }
      }
    };
    pass(false);
    pass(true);
    bool clean = true;
    for (size_t g = 0; g < kGuard; g++)
      clean = clean && shadow[2 * 1024 + g] == -7.0 && ovf[2 * (kPlaneOvfBytes / 8) + g] == -7.0 && implane[2 * kPlaneDoubles + g] == -7.0 && plane[2 * kPlaneDoubles + g] == -7.0;
    // the aliased planes stay inside their blocks: rows 0..29 end at double 30*34 = 1020 <= 1024, the overflow rows at (96 + 2*272) / 8 = 80
    return clean;
  };
  if (!run(0, ref) || !run(layout, got)) return -1;
  int bad = 0;
  for (int w = 0; w < 2; w++)
    for (int lane = 0; lane < 32; lane++)
      bad += memcmp(&ref[w][lane], &got[w][lane], sizeof(Regs)) != 0;
  return bad;
}
