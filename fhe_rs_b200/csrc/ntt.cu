// NTT launcher: picks the (N1, N2) split and tile shapes for a given N and instantiates
// the tile kernels of ntt.cuh.
#include <cstdlib>

#include "engine.hpp"
#include "ntt_fast.cuh"

namespace fhe_b200 {

std::atomic<unsigned long long> g_launches{0};

namespace {

constexpr int kTileLog = 12;  // 4096 words (32 KiB + padding) of shared memory per CTA

template <int LOGP, int LOGB, bool INV>
void run_rows(const NttArgs& a, cudaStream_t st) {
  constexpr u32 T = 1u << (LOGP + LOGB);
  const u32 tiles = (1u << a.logn1) >> LOGB;
  const u32 threads = T / 8 >= 256 ? 256 : (T / 8 < 32 ? 32 : T / 8);
  const size_t smem = (T + (T >> 5) + 1) * sizeof(u64);
  ntt_rows_kernel<LOGP, LOGB, INV><<<a.n_rows * tiles, threads, smem, st>>>(a);
  g_launches++;
}
template <int LOGP, int LOGB, bool INV>
void run_cols(const NttArgs& a, cudaStream_t st) {
  constexpr u32 T = 1u << (LOGP + LOGB);
  const u32 tiles = (1u << (a.logn - LOGP)) >> LOGB;
  const size_t smem = (T + (T >> 5) + 1) * sizeof(u64);
  ntt_cols_kernel<LOGP, LOGB, INV><<<a.n_rows * tiles, 256, smem, st>>>(a);
  g_launches++;
}

template <int LOGP, bool COLS, bool INV, int TLOG = 12>
void run_fast(const NttArgs& a, cudaStream_t st) {
  constexpr size_t smem = 2 * FastTile<LOGP, COLS, INV, false, TLOG>::TW * sizeof(u64);
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaFuncSetAttribute(ntt_fast_kernel<LOGP, COLS, INV, TLOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  constexpr int LOGB = TLOG - LOGP;
  const u32 tiles = COLS ? ((1u << (a.logn - LOGP)) >> LOGB) : ((1u << a.logn1) >> LOGB);
  ntt_fast_kernel<LOGP, COLS, INV, TLOG><<<a.n_rows * tiles, 1 << (TLOG - 3), smem, st>>>(a);
  g_launches++;
}
template <bool INV, int TLOG = 12>
void run_fast_cols_for(const NttArgs& a, cudaStream_t st) {
  switch (a.logn1) {
    case 7: run_fast<7, true, INV, TLOG>(a, st); break;
    case 8: run_fast<8, true, INV, TLOG>(a, st); break;
    case 9: run_fast<9, true, INV, TLOG>(a, st); break;
    case 10: run_fast<10, true, INV, TLOG>(a, st); break;
    default: break;
  }
}

template <bool INV>
void run_single(const NttArgs& a, cudaStream_t st) {
  switch (a.logn) {
    case 3: run_rows<3, 0, INV>(a, st); break;
    case 4: run_rows<4, 0, INV>(a, st); break;
    case 5: run_rows<5, 0, INV>(a, st); break;
    case 6: run_rows<6, 0, INV>(a, st); break;
    case 7: run_rows<7, 0, INV>(a, st); break;
    case 8: run_rows<8, 0, INV>(a, st); break;
    case 9: run_rows<9, 0, INV>(a, st); break;
    case 10: run_rows<10, 0, INV>(a, st); break;
    case 11: run_rows<11, 0, INV>(a, st); break;
    case 12: run_rows<12, 0, INV>(a, st); break;
    default: break;
  }
}
template <bool INV>
void run_cols_for(const NttArgs& a, cudaStream_t st) {
  switch (a.logn1) {
    case 7: run_cols<7, kTileLog - 7, INV>(a, st); break;
    case 8: run_cols<8, kTileLog - 8, INV>(a, st); break;
    case 9: run_cols<9, kTileLog - 9, INV>(a, st); break;
    case 10: run_cols<10, kTileLog - 10, INV>(a, st); break;
    default: break;
  }
}

}  // namespace

void launch_ntt(const u64* in, u64* out, u32 n_rows, const RowIds& ids, const LimbDev* limbs, u32 logn,
                bool inverse, u32 in_div, bool reduce_on_load, cudaStream_t st, bool lazy_out) {
  if (n_rows == 0) return;
  NttArgs a;
  a.in = in;
  a.out = out;
  a.limbs = limbs;
  a.n_rows = n_rows;
  a.limbs_per_poly = ids.limbs_per_poly;
  a.in_div = in_div;
  a.reduce_on_load = reduce_on_load ? 1 : 0;
  a.lazy_out = (lazy_out && !inverse) ? 1 : 0;
  a.logn = logn;
  for (int i = 0; i < kMaxPos; i++) a.ids[i] = ids.ids[i];
  if (logn <= 12) {
    a.logn1 = 0;
    if (inverse) run_single<true>(a, st); else run_single<false>(a, st);
    return;
  }
  // two passes: N2 = 64 contiguous points (rows kernel), N1 = N / 64 (cols kernel)
  a.logn1 = logn - 6;
  NttArgs second = a;  // second pass runs in place on `out`
  second.in = out;
  second.in_div = 1;
  second.reduce_on_load = 0;
  // the register-resident kernels carry the Shoup butterflies only (the Solinas form measured no faster and doubled
  // their code size); FHE_B200_SOLINAS_NTT therefore selects the generic tile kernels, which keep both
  static const bool generic = getenv("FHE_B200_GENERIC_NTT") != nullptr || getenv("FHE_B200_SOLINAS_NTT") != nullptr;
  if (generic) {
    if (!inverse) {
      run_cols_for<false>(a, st);
      run_rows<6, 6, false>(second, st);
    } else {
      run_rows<6, 6, true>(a, st);
      run_cols_for<true>(second, st);
    }
    return;
  }
  // Tile sizes.  Smaller CTAs (same 8 words and <= 64 registers per thread, so the same number of resident warps)
  // put more independent CTAs on an SM; their load / butterfly / exchange phases interleave and the multiplier
  // pipe idles less: measured at set C, 4096-word tiles for both passes 3590 products/s, rows pass 2048 words 3710,
  // 1024 words 3810 (512 words: no further gain), + cols pass 2048 words 3835 (1024 words, i.e. 16-byte column
  // segments at N = 2^15: 3600).  FHE_B200_ROWS_TLOG / FHE_B200_COLS_TLOG = 12 select the 4096-word tiles.
  static const int rows_tlog = getenv("FHE_B200_ROWS_TLOG") ? atoi(getenv("FHE_B200_ROWS_TLOG")) : 10;
  static const int cols_tlog = getenv("FHE_B200_COLS_TLOG") ? atoi(getenv("FHE_B200_COLS_TLOG")) : 11;
  auto rows = [&](const NttArgs& x, bool inv) {
    if (rows_tlog == 10) { if (inv) run_fast<6, false, true, 10>(x, st); else run_fast<6, false, false, 10>(x, st); }
    else { if (inv) run_fast<6, false, true>(x, st); else run_fast<6, false, false>(x, st); }
  };
  auto cols = [&](const NttArgs& x, bool inv) {
    // (N = 2^16 would leave a 2048-word tile two columns wide: keep the 4096-word tile there)
    if (cols_tlog == 11 && x.logn1 <= 9) { if (inv) run_fast_cols_for<true, 11>(x, st); else run_fast_cols_for<false, 11>(x, st); }
    else { if (inv) run_fast_cols_for<true>(x, st); else run_fast_cols_for<false>(x, st); }
  };
  if (!inverse) {
    cols(a, false);
    rows(second, false);
  } else {
    rows(a, true);
    cols(second, true);
  }
}

}  // namespace fhe_b200
