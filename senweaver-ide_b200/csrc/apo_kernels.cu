// apo_kernels.cu — sm_100a kernels of the APO scoring engine.
//
//   K1  k_reward9<36>     Form D (fp32 dims[C][T][9], 36 B/eval)  -> per-candidate exact partial sums
//   K1r k_reward9<32|16>  Form R / R16 records per (c,t) -> same, dims derived on device (TCS:668-763)
//   K2  corpus scan       records[T]: 6-pattern scan + tallies + per-dimension sums (apo_corpus.cuh):
//                         one extra warp of every K1 CTA, or the stand-alone k_detect6 launch
//   K3  finalize          segmented sum + radix top-K: last CTA of the scoring launch (1 rank), or
//                         k_finalize after the allreduce (> 1 rank)
//   k_gen_* / k_reward_batch: generators and the single-trace path
//   (K1q, the Form Q kernel, lives in apo_compact.cu)
//
// HBM-bound integer/fp64 streaming: no tensor cores on this path.  K1/K1r move tiles with 1-D
// TMA bulk copies (cp.async.bulk -> SASS UBLKCP) through an mbarrier full/empty ring fed by a
// dedicated producer warp; consumers read the staged tile with conflict-free LDS.128.
#include <cstdlib>
#include "apo_device.cuh"
#include "apo_format.h"
#include "apo_kernels.h"
#include "apo_corpus.cuh"

namespace apo {

// =================================================================== K1 / K1r
// Stage metadata written by the producer before it arms the full barrier.
struct StageMeta { int32_t cand; int32_t n; };   // n == 0 -> no more work

template <int ROW, int CW, int STAGES>
struct K1Cfg {
	static constexpr int EPT = 4;                       // evaluations per consumer thread per tile
	static constexpr int NCONS = CW * 32;
	static constexpr int TILE = NCONS * EPT;            // evaluations per tile
	static constexpr int STAGE_BYTES = TILE * ROW;
	static constexpr int LUT_BYTES = 512 * 16 + CAT_WORDS * 8;   // {total weight, reciprocal} per presence mask + categorical products
	static constexpr int BAR_OFF = STAGES * STAGE_BYTES + LUT_BYTES;
	static constexpr int META_OFF = BAR_OFF + 2 * STAGES * 8;
	static constexpr int EX_OFF = META_OFF + STAGES * 8;          // 18 example slots + the last-CTA flag of the fused corpus scan
	static constexpr int SMEM = EX_OFF + 18 * 8 + 16;
	static_assert(STAGE_BYTES % 16 == 0, "bulk copies are multiples of 16 bytes");
	// 227 KB per CTA minus the kernel's static shared memory (1408 B, ptxas -v): the 24-warp Form D tile sits 48 B under the limit
	static_assert(SMEM <= 232448 - 1408, "tile + tables exceed the shared memory of one CTA");
};

// One Form-D evaluation, first half: 9 fp32 (NaN = absent) -> weighted sum in push order and
// the LUT entry {total weight, reciprocal} of its presence mask (TCS:777-783).
__device__ __forceinline__ void eval_ws(const float (&v)[APO_NDIM], const Weights &W, const double2 *lut,
                                        double &ws_out, double2 &t_out, uint32_t &valid) {
	double ws = 0.0;
	uint32_t mask = 0;
#pragma unroll
	for (int i = 0; i < APO_NDIM; i++) {
		const float f = v[i];
		const bool p = (f == f);
		const float g = p ? f : 0.0f;                 // +0.0 * w leaves the running sum unchanged
		ws = __dadd_rn(ws, __dmul_rn((double)g, W.w[i]));
		mask |= (p ? 1u : 0u) << lut_bit(i);
	}
	ws_out = ws;
	t_out = lut[mask];                                // total weight 0 -> {-1, -1}: ws == 0 -> quotient 0, not counted
	valid = t_out.x > 0.0 ? 1u : 0u;                  // TCS:784: totalWeight > 0 ? ... : null
}

template <bool RECIP>
__device__ __forceinline__ long long eval_dims(const float (&v)[APO_NDIM], const Weights &W, const double2 *lut,
                                               uint32_t &valid) {
	double ws; double2 t;
	eval_ws(v, W, lut, ws, t, valid);
	return to_fx(div_lut<RECIP>(ws, t));              // TCS:784
}

// One Form-R evaluation, first half (TCS:668-783): record -> weighted sum + LUT entry.  The
// categorical product table sits right behind the 512-entry LUT in shared memory.
__device__ __forceinline__ void record_ws(const apo_record &r, const Weights &W, const double2 *lut,
                                          double &ws_out, double2 &t_out) {
	const double *cat = reinterpret_cast<const double *>(lut + 512);
	const uint32_t mask = record_ws_direct(r, W.w[2], cat, ws_out);
	t_out = lut[lut_index(mask)];
}

template <bool RECIP>
__device__ __forceinline__ long long eval_record(const apo_record &r, const Weights &W, const double2 *lut,
                                                 uint32_t &valid) {
	double ws; double2 t;
	record_ws(r, W, lut, ws, t);
	valid = ((r.flags & APO_F_VALID) && t.x > 0.0) ? 1u : 0u;
	return valid ? to_fx(div_lut<RECIP>(ws, t)) : 0ll;
}

template <int ROW, int CW, int STAGES, bool RECIP>
__global__ void __launch_bounds__((CW + 2) * 32, 1)
k_reward9(const K1Params P) {
	using Cfg = K1Cfg<ROW, CW, STAGES>;
	extern __shared__ __align__(128) uint8_t smem[];
	double2 *s_lut = reinterpret_cast<double2 *>(smem + STAGES * Cfg::STAGE_BYTES);
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::BAR_OFF);
	uint64_t *empty = full + STAGES;
	StageMeta *meta = reinterpret_cast<StageMeta *>(smem + Cfg::META_OFF);
	unsigned long long *s_ex = reinterpret_cast<unsigned long long *>(smem + Cfg::EX_OFF);
	bool *s_last = reinterpret_cast<bool *>(smem + Cfg::EX_OFF + 18 * 8);

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int i = tid; i < 512; i += blockDim.x) s_lut[lut_index(i)] = make_double2(P.lut[i], P.lut[512 + i]);
	for (int i = tid; i < CAT_WORDS; i += blockDim.x) reinterpret_cast<double *>(s_lut + 512)[i] = P.lut[1024 + i];
	if (tid == 0) {
		for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
		mbar_fence_init();
	}
	__syncthreads();

	// contiguous range of tiles for this CTA; tiles are enumerated candidate-major
	const uint64_t lo = P.total_tiles * blockIdx.x / gridDim.x;
	const uint64_t hi = P.total_tiles * (blockIdx.x + 1ull) / gridDim.x;

	if (warp == CW) {
		// ------------------------------------------------ producer warp (one elected lane)
		if (lane == 0) {
			const uint64_t pol = policy_evict_first();
			uint32_t it = 0;
			for (uint64_t tile = lo; tile <= hi; ++tile, ++it) {
				const int s = it % STAGES;
				const uint32_t ph = (it / STAGES) & 1u;
				mbar_wait(&empty[s], ph ^ 1u);
				if (tile == hi) {                       // terminator
					meta[s].cand = -1; meta[s].n = 0;
					mbar_arrive(&full[s]);
					break;
				}
				const uint32_t c = (uint32_t)(tile / P.tiles_per_cand);
				const uint64_t j = tile - (uint64_t)c * P.tiles_per_cand;
				const uint64_t e0 = j * Cfg::TILE;
				const uint64_t rem = P.T - e0;
				const uint32_t n = rem < (uint64_t)Cfg::TILE ? (uint32_t)rem : (uint32_t)Cfg::TILE;
				const uint32_t bytes = ((n + 3u) & ~3u) * ROW;
				meta[s].cand = (int32_t)c; meta[s].n = (int32_t)n;
				mbar_expect_tx(&full[s], bytes);
				uint8_t *dst = smem + s * Cfg::STAGE_BYTES;
				const uint8_t *gsrc = P.base + (uint64_t)c * P.pitch_bytes + e0 * ROW;
				if (P.tune == 0) bulk_g2s(dst, gsrc, bytes, &full[s], pol);
				else {
					const uint32_t parts = 1u << ((P.tune >> 1) & 7u);
					uint32_t chunk = (bytes / parts + 15u) & ~15u;
					if (chunk == 0) chunk = bytes;
					for (uint32_t off = 0; off < bytes; off += chunk) {
						const uint32_t nb = bytes - off < chunk ? bytes - off : chunk;
						if (P.tune & 1u) bulk_g2s_nohint(dst + off, gsrc + off, nb, &full[s]);
						else bulk_g2s(dst + off, gsrc + off, nb, &full[s], pol);
					}
				}
			}
		}
	} else if (warp == CW + 1) {
		// ------------------------------------------------ corpus warp: K2's scan rides on the spare issue slots
		if (P.corpus_on) {
			if (lane < 18) s_ex[lane] = 0ull;
			__syncwarp();
			corpus_scan_warp<true>(P.corpus, (uint64_t)blockIdx.x * 32, (uint64_t)gridDim.x * 32, s_ex,
			                       reinterpret_cast<const double *>(s_lut + 512), s_lut, lane);
		}
	} else {
	// ---------------------------------------------------- consumer warps
	Acc128 acc; acc.zero();
	uint32_t cnt = 0;
	int cur = -1;
	const Weights W = P.W;
	auto flush = [&](int c) {
		long long *dst = P.acc + (uint64_t)ACC_PER_CAND * c;
		flush_acc128(acc, dst, lane);
		const uint32_t n = warp_sum_u32(cnt);
		if (lane == 0 && n) atomicAdd((unsigned long long *)dst + 3, (unsigned long long)n);
		acc.zero(); cnt = 0;
	};

	for (uint32_t it = 0;; ++it) {
		const int s = it % STAGES;
		const uint32_t ph = (it / STAGES) & 1u;
		mbar_wait(&full[s], ph);
		const int c = meta[s].cand, n = meta[s].n;
		if (n == 0) break;
		if (c != cur) { if (cur >= 0) flush(cur); cur = c; }
		const uint8_t *st = smem + s * Cfg::STAGE_BYTES;
		if (ROW == 36) {
			// thread owns 4 consecutive evaluations = 144 B = 9 x LDS.128 (conflict-free: 144 B lane stride)
			const int e0 = tid * 4;
			const float4 *src = reinterpret_cast<const float4 *>(st + (size_t)e0 * 36);
			if (n == Cfg::TILE) {
				// full tile: branch-free, the four evaluations are independent chains the scheduler interleaves
				float f[36];
#pragma unroll
				for (int q = 0; q < 9; q++) {
					const float4 x = src[q];
					f[4 * q] = x.x; f[4 * q + 1] = x.y; f[4 * q + 2] = x.z; f[4 * q + 3] = x.w;
				}
				double ws4[4]; double2 t4[4];
#pragma unroll
				for (int k = 0; k < 4; k++) {
					float v[APO_NDIM];
#pragma unroll
					for (int i = 0; i < APO_NDIM; i++) v[i] = f[9 * k + i];
					uint32_t ok;
					eval_ws(v, W, s_lut, ws4[k], t4[k], ok);
					cnt += ok;
				}
				// one uniform check for the (rare) flagged masks keeps the four divisions branch-free
				const bool generic = !RECIP && ((__double2hiint(t4[0].y) | __double2hiint(t4[1].y) |
				                                 __double2hiint(t4[2].y) | __double2hiint(t4[3].y)) < 0);
				long long x4 = 0;                         // |fr| <= 512 -> |x| < 2^61: four terms fit int64
				if (!generic) {
#pragma unroll
					for (int k = 0; k < 4; k++) x4 += to_fx(div_fast<RECIP>(ws4[k], t4[k]));
				} else {
#pragma unroll
					for (int k = 0; k < 4; k++) x4 += to_fx(div_lut<RECIP>(ws4[k], t4[k]));
				}
				acc.add(x4);
			} else if (e0 < n) {
				float f[36];
#pragma unroll
				for (int q = 0; q < 9; q++) {
					const float4 x = src[q];
					f[4 * q] = x.x; f[4 * q + 1] = x.y; f[4 * q + 2] = x.z; f[4 * q + 3] = x.w;
				}
#pragma unroll
				for (int k = 0; k < 4; k++) {
					if (e0 + k < n) {
						float v[APO_NDIM];
#pragma unroll
						for (int i = 0; i < APO_NDIM; i++) v[i] = f[9 * k + i];
						uint32_t ok;
						const long long x = eval_dims<RECIP>(v, W, s_lut, ok);
						acc.add(x);
						cnt += ok;
					}
				}
			}
		} else {
			// Form R / R16: evaluation e = k*NCONS + tid, 32 B = 2 x LDS.128 or 16 B = 1 x LDS.128
			auto load_record = [&](int e) -> apo_record {
				if (ROW == 32) {
					const uint4 *src = reinterpret_cast<const uint4 *>(st + (size_t)e * 32);
					union { uint4 q[2]; apo_record r; } u;
					u.q[0] = src[0]; u.q[1] = src[1];
					return u.r;
				} else {
					union { uint4 q; apo_record16 p; } u;
					u.q = *reinterpret_cast<const uint4 *>(st + (size_t)e * 16);
					return unpack16(u.p);
				}
			};
			if (n == Cfg::TILE) {
				double ws4[Cfg::EPT]; double2 t4[Cfg::EPT]; uint32_t ok4[Cfg::EPT];
#pragma unroll
				for (int k = 0; k < Cfg::EPT; k++) {
					const apo_record r = load_record(k * Cfg::NCONS + tid);
					record_ws(r, W, s_lut, ws4[k], t4[k]);
					ok4[k] = ((r.flags & APO_F_VALID) && t4[k].x > 0.0) ? 1u : 0u;
					cnt += ok4[k];
				}
				const bool generic = !RECIP && ((__double2hiint(t4[0].y) | __double2hiint(t4[1].y) |
				                                 __double2hiint(t4[2].y) | __double2hiint(t4[3].y)) < 0);
				long long x4 = 0;
				if (!generic) {
#pragma unroll
					for (int k = 0; k < Cfg::EPT; k++) { const long long x = to_fx(div_fast<RECIP>(ws4[k], t4[k])); x4 += ok4[k] ? x : 0ll; }
				} else {
#pragma unroll
					for (int k = 0; k < Cfg::EPT; k++) { const long long x = to_fx(div_lut<RECIP>(ws4[k], t4[k])); x4 += ok4[k] ? x : 0ll; }
				}
				acc.add(x4);
			} else {
#pragma unroll
				for (int k = 0; k < Cfg::EPT; k++) {
					const int e = k * Cfg::NCONS + tid;
					if (e < n) {
						const apo_record r = load_record(e);
						uint32_t ok;
						const long long x = eval_record<RECIP>(r, W, s_lut, ok);
						acc.add(x);
						cnt += ok;
					}
				}
			}
		}
		__syncwarp();
		if (lane == 0) mbar_arrive(&empty[s]);
	}
	if (cur >= 0) flush(cur);
	}
	// block-wide: every warp arrives here; the stage buffers are free by now (every tile was consumed) and serve as scratch
	if (P.corpus_on) corpus_tail(P.corpus, s_ex, s_last, reinterpret_cast<unsigned long long *>(smem), 2048u);
}

template <int ROW, int CW, int STAGES>
static cudaError_t launch_k1(const K1Params &P, int grid, bool recip, cudaStream_t st) {
	using Cfg = K1Cfg<ROW, CW, STAGES>;
	cudaError_t err;
	if (recip) {
		auto k = k_reward9<ROW, CW, STAGES, true>;
		if ((err = allow_big_smem(k, Cfg::SMEM)) != cudaSuccess) return err;
		k<<<grid, (CW + 2) * 32, Cfg::SMEM, st>>>(P);
	} else {
		auto k = k_reward9<ROW, CW, STAGES, false>;
		if ((err = allow_big_smem(k, Cfg::SMEM)) != cudaSuccess) return err;
		k<<<grid, (CW + 2) * 32, Cfg::SMEM, st>>>(P);
	}
	return cudaGetLastError();
}

int k1_tile_evals(int row, int variant) {
	(void)row;
	// variant 0 (default) = 20 consumer warps x 2 stages of 92 KB: measured best on B200
	// (12.15 ms at 256 x 10M vs 13.8 ms for 8 warps x 5 stages) — warps hide the fp64 chains,
	// two 92 KB tiles in flight per SM already cover the HBM latency-bandwidth product.
	switch (variant) {
	case 1: return K1Cfg<36, 16, 3>::TILE;
	case 2: return K1Cfg<36, 12, 4>::TILE;
	case 3: return K1Cfg<36, 8, 5>::TILE;
	case 4: return K1Cfg<36, 24, 2>::TILE;
	default: return K1Cfg<36, 20, 2>::TILE;
	}
}

cudaError_t run_reward9(K1Params P, int row, int variant, bool recip, int sm_count, cudaStream_t st) {
	const int tile = k1_tile_evals(row, variant);
	P.tiles_per_cand = (uint32_t)((P.T + tile - 1) / tile);
	P.total_tiles = (uint64_t)P.tiles_per_cand * P.C;
	if (P.total_tiles == 0) return cudaSuccess;
	int grid = sm_count;
	{ static const int env_grid = [] { const char *g = getenv("APO_K1_GRID"); return g ? atoi(g) : 0; }(); if (env_grid > 0 && env_grid < grid) grid = env_grid; }   // tuning experiments only, read once
	{ static const int env_tune = [] { const char *g = getenv("APO_K1_TUNE"); return g ? atoi(g) : 0; }(); if (env_tune) P.tune = (uint32_t)env_tune; }
	if ((uint64_t)grid > P.total_tiles) grid = (int)P.total_tiles;
	if (P.corpus_on && P.corpus.T) {
		// a tiny evaluation set with a corpus to scan: CTAs without tiles still lend their corpus warp (<= 2 records per lane)
		const uint64_t want = (P.corpus.T + 63) / 64;
		const int more = (int)(want < (uint64_t)sm_count ? want : (uint64_t)sm_count);
		if (more > grid) grid = more;
	}
	if (row == 36) {
		switch (variant) {
		case 1: return launch_k1<36, 16, 3>(P, grid, recip, st);
		case 2: return launch_k1<36, 12, 4>(P, grid, recip, st);
		case 3: return launch_k1<36, 8, 5>(P, grid, recip, st);
		case 4: return launch_k1<36, 24, 2>(P, grid, recip, st);
		default: return launch_k1<36, 20, 2>(P, grid, recip, st);
		}
	} else if (row == 16) {
		switch (variant) {
		case 1: return launch_k1<16, 16, 3>(P, grid, recip, st);
		case 2: return launch_k1<16, 12, 4>(P, grid, recip, st);
		case 3: return launch_k1<16, 8, 5>(P, grid, recip, st);
		case 4: return launch_k1<16, 24, 2>(P, grid, recip, st);
		default: return launch_k1<16, 20, 2>(P, grid, recip, st);
		}
	} else {
		switch (variant) {
		case 1: return launch_k1<32, 16, 3>(P, grid, recip, st);
		case 2: return launch_k1<32, 12, 4>(P, grid, recip, st);
		case 3: return launch_k1<32, 8, 5>(P, grid, recip, st);
		case 4: return launch_k1<32, 24, 2>(P, grid, recip, st);
		default: return launch_k1<32, 20, 2>(P, grid, recip, st);
		}
	}
}

// =================================================================== K2 detect6 / K3 finalize (stand-alone launches)
constexpr int K2_THREADS = 256;

__global__ void __launch_bounds__(K2_THREADS, 2)
k_detect6(const K2Params P) {
	__shared__ unsigned long long s_ex[APO_NPAT * 3];
	__shared__ double s_cat[CAT_WORDS];
	__shared__ double2 s_lut[512];                     // {total weight, reciprocal} per (natural) presence mask
	__shared__ bool s_last;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid < APO_NPAT * 3) s_ex[tid] = 0ull;
	if (tid < CAT_WORDS) s_cat[tid] = P.lut[1024 + tid];
	for (int i = tid; i < 512; i += K2_THREADS) s_lut[i] = make_double2(P.lut[i], P.lut[512 + i]);
	__syncthreads();
	const uint64_t nwarps = (uint64_t)gridDim.x * (K2_THREADS / 32);
	corpus_scan_warp<false>(P, ((uint64_t)blockIdx.x * (K2_THREADS / 32) + warp) * 32, nwarps * 32, s_ex, s_cat, s_lut, lane);
	corpus_tail(P, s_ex, &s_last, reinterpret_cast<unsigned long long *>(s_lut), 1024u);     // the LUT (8 KB) is done with: scratch of the tail
}

// The stand-alone corpus scan as a streaming kernel (what sessions, host streaming and small calls pay when the scan cannot
// ride inside a scoring launch): one persistent CTA per SM, a producer warp moving 64 KB record tiles with 1-D TMA bulk
// copies through a 3-stage mbarrier ring, 16 consumer warps scoring 4 records per thread per tile out of shared memory —
// the grid-stride form above keeps one 32-byte record per thread in flight (16 KB per SM, a third of what the HBM latency x
// bandwidth product needs) and stalls on every load.
template <int CW, int RPT, int STAGES>
struct K2Cfg {
	static constexpr int NCONS = CW * 32;
	static constexpr int TILE = NCONS * RPT;                 // records per tile
	static constexpr int STAGE_BYTES = TILE * 32;
	static constexpr int LUT_OFF = STAGES * STAGE_BYTES;     // double2[512]
	static constexpr int CAT_OFF = LUT_OFF + 512 * 16;
	static constexpr int BAR_OFF = CAT_OFF + CAT_WORDS * 8;
	static constexpr int META_OFF = BAR_OFF + 2 * STAGES * 8;
	static constexpr int EX_OFF = META_OFF + STAGES * 8;
	static constexpr int SMEM = EX_OFF + 18 * 8 + 16;
	static_assert(SMEM <= 232448 - 1408, "tiles + tables exceed the shared memory of one CTA");
};

template <int CW, int RPT, int STAGES>
__global__ void __launch_bounds__((CW + 1) * 32, 1)
k_detect6_tiles(const K2Params P, uint64_t total_tiles) {
	using Cfg = K2Cfg<CW, RPT, STAGES>;
	extern __shared__ __align__(128) uint8_t smem[];
	double2 *s_lut = reinterpret_cast<double2 *>(smem + Cfg::LUT_OFF);
	double *s_cat = reinterpret_cast<double *>(smem + Cfg::CAT_OFF);
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::BAR_OFF);
	uint64_t *empty = full + STAGES;
	StageMeta *meta = reinterpret_cast<StageMeta *>(smem + Cfg::META_OFF);
	unsigned long long *s_ex = reinterpret_cast<unsigned long long *>(smem + Cfg::EX_OFF);
	bool *s_last = reinterpret_cast<bool *>(smem + Cfg::EX_OFF + 18 * 8);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	for (int i = tid; i < 512; i += blockDim.x) s_lut[i] = make_double2(P.lut[i], P.lut[512 + i]);
	for (int i = tid; i < CAT_WORDS; i += blockDim.x) s_cat[i] = P.lut[1024 + i];
	if (tid < 18) s_ex[tid] = 0ull;
	if (tid == 0) {
		for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
		mbar_fence_init();
	}
	__syncthreads();
	const uint64_t lo = total_tiles * blockIdx.x / gridDim.x;
	const uint64_t hi = total_tiles * (blockIdx.x + 1ull) / gridDim.x;

	if (warp == CW) {
		if (lane == 0) {
			const uint64_t pol = policy_evict_first();
			uint32_t it = 0;
			for (uint64_t tile = lo; tile <= hi; ++tile, ++it) {
				const int s = it % STAGES;
				const uint32_t ph = (it / STAGES) & 1u;
				mbar_wait(&empty[s], ph ^ 1u);
				if (tile == hi) { meta[s].cand = -1; meta[s].n = 0; mbar_arrive(&full[s]); break; }
				const uint64_t e0 = tile * Cfg::TILE;
				const uint64_t rem = P.T - e0;
				const uint32_t n = rem < (uint64_t)Cfg::TILE ? (uint32_t)rem : (uint32_t)Cfg::TILE;
				meta[s].cand = 0; meta[s].n = (int32_t)n;
				mbar_expect_tx(&full[s], n * 32u);
				bulk_g2s(smem + s * Cfg::STAGE_BYTES, reinterpret_cast<const uint8_t *>(P.recs) + e0 * 32, n * 32u, &full[s], pol);
			}
		}
	} else {
		const double w2 = P.W.w[2];
		long long *corp = P.acc + (uint64_t)ACC_PER_CAND * P.C;
		ScanAcc A; A.zero();
		int since = 0;                                            // tiles since the last flush: RPT records per thread each
		uint64_t tile = lo;
		for (uint32_t it = 0;; ++it, ++tile) {
			const int s = it % STAGES;
			const uint32_t ph = (it / STAGES) & 1u;
			mbar_wait(&full[s], ph);
			const int n = meta[s].n;
			if (n == 0) break;
			const uint8_t *st = smem + s * Cfg::STAGE_BYTES;
			const unsigned long long g0 = P.idx_base + tile * Cfg::TILE;
#pragma unroll
			for (int k = 0; k < RPT; k++) {
				const int e = k * Cfg::NCONS + tid;
				if (e < n) {
					const uint4 *src = reinterpret_cast<const uint4 *>(st + (size_t)e * 32);
					union { uint4 q[2]; apo_record r; } u;
					u.q[0] = src[0]; u.q[1] = src[1];
					scan_record<false>(A, u.r, g0 + (unsigned long long)e, s_ex, s_cat, s_lut, w2);
				}
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(&empty[s]);
			if (++since == K2_CHUNK / RPT) { scan_flush(A, corp, lane); A.zero(); since = 0; }
		}
		if (since) scan_flush(A, corp, lane);
	}
	corpus_tail(P, s_ex, s_last, reinterpret_cast<unsigned long long *>(smem), 2048u);     // block-wide: every warp arrives here
}

template <int CW, int RPT, int STAGES>
static cudaError_t launch_k2_tiles(const K2Params &P, int sm_count, cudaStream_t st) {
	using Cfg = K2Cfg<CW, RPT, STAGES>;
	const uint64_t tiles = (P.T + Cfg::TILE - 1) / Cfg::TILE;
	auto k = k_detect6_tiles<CW, RPT, STAGES>;
	cudaError_t err = allow_big_smem(k, Cfg::SMEM);
	if (err != cudaSuccess) return err;
	int grid = sm_count;
	if ((uint64_t)grid > tiles) grid = (int)tiles;
	k<<<grid, (CW + 1) * 32, Cfg::SMEM, st>>>(P, tiles);
	return cudaGetLastError();
}

cudaError_t run_detect6(const K2Params &P, int sm_count, cudaStream_t st) {
	static const int env_old = [] { const char *g = getenv("APO_K2_GRIDSTRIDE"); return g ? atoi(g) : 0; }();   // A/B against the grid-stride form
	static const int env_cw = [] { const char *g = getenv("APO_K2_CW"); return g ? atoi(g) : 0; }();            // A/B: consumer warps per CTA
	if (P.T >= 4 * 2048 && !env_old) {
		if (env_cw == 24) return launch_k2_tiles<24, 2, 4>(P, sm_count, st);
		if (env_cw == 20) return launch_k2_tiles<20, 2, 4>(P, sm_count, st);
		return launch_k2_tiles<16, 4, 3>(P, sm_count, st);      // measured: 16 warps x 4 records 0.225 ms / 10 M records, 20 x 2: 0.252, 24 x 2: 0.306 (register spills)
	}
	// a handful of records (the IDE's real corpora: <= 1000 traces): the light grid-stride kernel, no 200 KB of shared memory to set up
	uint64_t want = (P.T + K2_THREADS - 1) / K2_THREADS;
	int grid = sm_count * 2;            // two resident CTAs per SM (launch bounds), one wave
	if ((uint64_t)grid > want) grid = (int)(want ? want : 1);
	k_detect6<<<grid, K2_THREADS, 0, st>>>(P);
	return cudaGetLastError();
}

__global__ void __launch_bounds__(1024) k_finalize(const FinalizeParams F) {
	__shared__ unsigned long long s_keys[1024];
	finalize_and_publish(F, s_keys, 1024u);
}

cudaError_t run_finalize(const FinalizeParams &F, cudaStream_t st) {
	k_finalize<<<1, 1024, 0, st>>>(F);
	return cudaGetLastError();
}

// =================================================================== generators / single-trace path
__global__ void __launch_bounds__(256)
k_gen_dims(float *out, uint64_t pitch_evals, unsigned long long seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T,
           uint32_t agent_permille) {
	const uint32_t c = blockIdx.y;
	if (c >= C) return;
	const unsigned long long key = gen_key(seed, STREAM_ROLLOUT, c0 + c);
	const uint32_t qc = gen_quality(seed, STREAM_ROLLOUT, c0 + c);
	const float qnan = __int_as_float(0x7fc00000);
	for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x) {
		const apo_record r = gen_record(key, qc, t0 + t, agent_permille);
		double d[APO_NDIM];
		const uint32_t mask = reward_dims(r, d);
		float *row = out + ((uint64_t)c * pitch_evals + t) * APO_NDIM;
		const bool valid = (r.flags & APO_F_VALID) != 0;
#pragma unroll
		for (int i = 0; i < APO_NDIM; i++) row[i] = (valid && ((mask >> i) & 1u)) ? __double2float_rn(d[i]) : qnan;
	}
}

__global__ void __launch_bounds__(256)
k_gen_records(apo_record *out, uint64_t pitch, unsigned long long seed, uint32_t stream, uint32_t c0, uint32_t C,
              uint64_t t0, uint64_t T, uint32_t agent_permille) {
	const uint32_t c = blockIdx.y;
	if (c >= C) return;
	const unsigned long long key = gen_key(seed, stream, c0 + c);
	const uint32_t qc = gen_quality(seed, stream, c0 + c);
	for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x) {
		union { uint4 q[2]; apo_record r; } u;
		u.r = gen_record(key, qc, t0 + t, agent_permille);
		uint4 *dst = reinterpret_cast<uint4 *>(out + (uint64_t)c * pitch + t);
		dst[0] = u.q[0]; dst[1] = u.q[1];
	}
}

__global__ void __launch_bounds__(256)
k_gen_records16(apo_record16 *out, uint64_t pitch, unsigned long long seed, uint32_t stream, uint32_t c0, uint32_t C,
                uint64_t t0, uint64_t T, uint32_t agent_permille) {
	const uint32_t c = blockIdx.y;
	if (c >= C) return;
	const unsigned long long key = gen_key(seed, stream, c0 + c);
	const uint32_t qc = gen_quality(seed, stream, c0 + c);
	for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x) {
		union { uint4 q; apo_record16 p; } u;
		u.p = pack16(gen_record(key, qc, t0 + t, agent_permille));      // generator records are always representable
		*reinterpret_cast<uint4 *>(out + (uint64_t)c * pitch + t) = u.q;
	}
}

cudaError_t run_gen_records16(apo_record16 *out, uint64_t pitch, uint64_t seed, uint32_t stream, uint32_t c0, uint32_t C,
                              uint64_t t0, uint64_t T, uint32_t agent_permille, cudaStream_t st) {
	if (C == 0 || T == 0) return cudaSuccess;
	uint64_t gx = (T + 255) / 256;
	if (gx > 148ull * 64) gx = 148ull * 64;
	for (uint32_t cb = 0; cb < C; cb += 32768) {
		const uint32_t cn = C - cb < 32768 ? C - cb : 32768;
		k_gen_records16<<<dim3((unsigned)gx, cn), 256, 0, st>>>(out + (uint64_t)cb * pitch, pitch, seed, stream, c0 + cb, cn, t0, T, agent_permille);
	}
	return cudaGetLastError();
}

cudaError_t run_gen_dims(float *out, uint64_t pitch_evals, uint64_t seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T,
                         uint32_t agent_permille, cudaStream_t st) {
	if (C == 0 || T == 0) return cudaSuccess;
	uint64_t gx = (T + 255) / 256;
	if (gx > 148ull * 64) gx = 148ull * 64;
	for (uint32_t cb = 0; cb < C; cb += 32768) {       // gridDim.y limit 65535
		const uint32_t cn = C - cb < 32768 ? C - cb : 32768;
		k_gen_dims<<<dim3((unsigned)gx, cn), 256, 0, st>>>(out + (uint64_t)cb * pitch_evals * APO_NDIM, pitch_evals, seed, c0 + cb, cn, t0, T,
		                                                  agent_permille);
	}
	return cudaGetLastError();
}

cudaError_t run_gen_records(apo_record *out, uint64_t pitch, uint64_t seed, uint32_t stream, uint32_t c0, uint32_t C,
                            uint64_t t0, uint64_t T, uint32_t agent_permille, cudaStream_t st) {
	if (C == 0 || T == 0) return cudaSuccess;
	uint64_t gx = (T + 255) / 256;
	if (gx > 148ull * 64) gx = 148ull * 64;
	for (uint32_t cb = 0; cb < C; cb += 32768) {
		const uint32_t cn = C - cb < 32768 ? C - cb : 32768;
		k_gen_records<<<dim3((unsigned)gx, cn), 256, 0, st>>>(out + (uint64_t)cb * pitch, pitch, seed, stream, c0 + cb, cn, t0, T, agent_permille);
	}
	return cudaGetLastError();
}

// Reads a peer-mapped block once (apo_comm_init): the first access to lazily enabled peer memory pays for the mapping — 0.2 ms when
// it happened inside the first joined call.
__global__ void __launch_bounds__(256)
k_touch(const long long *p, uint64_t n, long long *sink) {
	long long s = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) s += __ldcg(p + i);
	if (s == 0x7fffffffffffffffll) *sink = s;                    // never true for zero-filled blocks: keeps the loads alive
}

cudaError_t run_touch(const void *p, uint64_t bytes, void *sink, cudaStream_t st) {
	if (bytes < 8) return cudaSuccess;
	k_touch<<<64, 256, 0, st>>>(reinterpret_cast<const long long *>(p), bytes / 8, reinterpret_cast<long long *>(sink));
	return cudaGetLastError();
}

// Records that arrive without the duration class (durClass == 0: hand-built records, older producers) get it here, once, from the
// binary32 duration — the comparisons the streaming kernels would otherwise repeat per evaluation.  rows = n records of 32 or 16 bytes.
__global__ void __launch_bounds__(256)
k_fill_durclass(uint8_t *rows, uint32_t row_bytes, uint64_t n) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint8_t *p = rows + i * row_bytes;
		if (row_bytes == 32) {
			apo_record *r = reinterpret_cast<apo_record *>(p);
			if (r->durClass & APO_DC_SET) continue;
			const double total = (double)(r->toolCalls ? r->toolCalls : 1u);
			r->durClass = (uint8_t)dur_class_from_float((double)r->toolDurMs, total, r->toolCalls > 0);
		} else {
			apo_record16 *r = reinterpret_cast<apo_record16 *>(p);
			if (r->durClass & APO_DC_SET) continue;
			const double total = (double)(r->toolCalls ? r->toolCalls : 1u);
			r->durClass = (uint8_t)dur_class_from_float((double)r->toolDurMs, total, r->toolCalls > 0);
		}
	}
}

cudaError_t run_fill_durclass(void *rows, uint32_t row_bytes, uint64_t n, cudaStream_t st) {
	if (n == 0) return cudaSuccess;
	uint64_t g = (n + 255) / 256;
	if (g > 148ull * 16) g = 148ull * 16;
	k_fill_durclass<<<(unsigned)g, 256, 0, st>>>(reinterpret_cast<uint8_t *>(rows), row_bytes, n);
	return cudaGetLastError();
}

// TraceCollectorService._computeRewardSignals for a batch of records (TCS:668-788).
__global__ void __launch_bounds__(256)
k_reward_batch(const apo_record *recs, uint64_t n, const Weights W, const double *lut, double *dims, uint32_t *masks,
               double *finals) {
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const apo_record r = recs[i];
	double d[APO_NDIM];
	const uint32_t mask = reward_dims(r, d);
	const double tw = lut[mask];
	const double fr = final_reward<false>(d, mask, W, lut);
	const double qnan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
	for (int k = 0; k < APO_NDIM; k++) dims[i * APO_NDIM + k] = ((mask >> k) & 1u) ? d[k] : qnan;
	masks[i] = mask;
	finals[i] = ((r.flags & APO_F_VALID) && tw > 0.0) ? fr : qnan;
}

cudaError_t run_reward_batch(const apo_record *recs, uint64_t n, const Weights &W, const double *lut, double *dims,
                             uint32_t *masks, double *finals, cudaStream_t st) {
	if (n == 0) return cudaSuccess;
	k_reward_batch<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(recs, n, W, lut, dims, masks, finals);
	return cudaGetLastError();
}

}  // namespace apo
