// apo_kernels.cu — sm_100a kernels of the APO scoring engine.
//
//   K1  reward9        Form D (fp32 dims[C][T][9], 36 B/eval)  -> per-candidate exact partial sums
//   K1r reward9_raw    Form R (apo_record per (c,t), 32 B/eval) -> same, dims derived on device
//   K2  detect6        corpus records[T]: 6-pattern scan + tallies + per-dimension sums
//                      (+ fused segmented sum / radix top-K in the last CTA at 1 rank)
//   K3  finalize       segmented sum + radix top-K after the allreduce (> 1 rank)
//   gen_* / reward_batch: generators and the single-trace path
//
// All of them are HBM-bound integer/fp64 streaming kernels: no tensor cores on this path.
// K1/K1r move tiles with 1-D TMA bulk copies (cp.async.bulk -> SASS UBLKCP) through an
// mbarrier full/empty ring fed by a dedicated producer warp; consumers read the staged
// tile with conflict-free LDS.128.
#include <cstdlib>
#include "apo_device.cuh"
#include "apo_format.h"
#include "apo_kernels.h"

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
	static constexpr int SMEM = META_OFF + STAGES * 8;
	static_assert(STAGE_BYTES % 16 == 0, "bulk copies are multiples of 16 bytes");
};

// One Form-D evaluation: 9 fp32 (NaN = absent) -> fixed-point finalReward.
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
	const uint32_t mask = record_ws_table(r, W.w[2], cat, ws_out);
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
__global__ void __launch_bounds__((CW + 1) * 32, 1)
k_reward9(const K1Params P) {
	using Cfg = K1Cfg<ROW, CW, STAGES>;
	extern __shared__ __align__(128) uint8_t smem[];
	double2 *s_lut = reinterpret_cast<double2 *>(smem + STAGES * Cfg::STAGE_BYTES);
	uint64_t *full = reinterpret_cast<uint64_t *>(smem + Cfg::BAR_OFF);
	uint64_t *empty = full + STAGES;
	StageMeta *meta = reinterpret_cast<StageMeta *>(smem + Cfg::META_OFF);

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
		return;
	}

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

template <int ROW, int CW, int STAGES>
static cudaError_t launch_k1(const K1Params &P, int grid, bool recip, cudaStream_t st) {
	using Cfg = K1Cfg<ROW, CW, STAGES>;
	cudaError_t err;
	if (recip) {
		auto k = k_reward9<ROW, CW, STAGES, true>;
		if ((err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)) != cudaSuccess) return err;
		k<<<grid, (CW + 1) * 32, Cfg::SMEM, st>>>(P);
	} else {
		auto k = k_reward9<ROW, CW, STAGES, false>;
		if ((err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM)) != cudaSuccess) return err;
		k<<<grid, (CW + 1) * 32, Cfg::SMEM, st>>>(P);
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
	if (const char *g = getenv("APO_K1_GRID")) { const int v = atoi(g); if (v > 0 && v < grid) grid = v; }   // tuning experiments only
	if (const char *g = getenv("APO_K1_TUNE")) P.tune = (uint32_t)atoi(g);
	if ((uint64_t)grid > P.total_tiles) grid = (int)P.total_tiles;
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

// =================================================================== K2 detect6
constexpr int K2_THREADS = 256;
__device__ __forceinline__ void ex_insert(unsigned long long *slots, unsigned long long idx) {
	// keep the 3 smallest indices: a value flows down the chain, every slot only decreases
	unsigned long long v = idx;
#pragma unroll
	for (int k = 0; k < 3; k++) {
		const unsigned long long old = atomicMin(&slots[k], v);
		if (old == ~0ull) return;                   // slot was empty: nothing displaced
		v = old > v ? old : v;
	}
}

__device__ void finalize_block(const FinalizeParams &F);

// Per-thread state of one chunk of <= K2_CHUNK records.  Seven of the nine dimensions are
// categorical (TCS:677-761), so their per-dimension sums (APO:556-565) are kept as packed
// occurrence counters — one 64-bit register per dimension, one small field per category — and
// turned into exact fixed-point sums at flush time: sum = SUM_k count_k * rint(value_k * 2^52),
// an integer product, identical to adding rint(value * 2^52) record by record.
constexpr int K2_CHUNK = 256;      // < 2^9: no packed field (>= 9 bits wide) can overflow inside a chunk

__device__ __forceinline__ void add_i128(long long *dst, __int128 v) {
	const unsigned long long lo = (unsigned long long)v;
	const unsigned long long l0 = lo & 0xffffffffull, l1 = lo >> 32, l2 = (unsigned long long)(long long)(v >> 64);
	if (l0) atomicAdd((unsigned long long *)dst + 0, l0);
	if (l1) atomicAdd((unsigned long long *)dst + 1, l1);
	if (l2) atomicAdd((unsigned long long *)dst + 2, l2);
}
// field f (width bits) of a packed counter, summed over the warp
template <int WIDTH>
__device__ __forceinline__ uint32_t field_sum(unsigned long long pk, int f) {
	return warp_sum_u32((uint32_t)(pk >> (WIDTH * f)) & ((1u << WIDTH) - 1u));
}

__global__ void __launch_bounds__(K2_THREADS, 2)
k_detect6(const K2Params P) {
	__shared__ unsigned long long s_ex[APO_NPAT * 3];
	__shared__ double s_cat[CAT_WORDS];
	__shared__ double2 s_lut[512];                     // {total weight, reciprocal} per (natural) presence mask
	__shared__ bool s_last;
	const int tid = threadIdx.x, lane = tid & 31;
	if (tid < APO_NPAT * 3) s_ex[tid] = ~0ull;
	if (tid < CAT_WORDS) s_cat[tid] = P.lut[1024 + tid];
	for (int i = tid; i < 512; i += K2_THREADS) s_lut[i] = make_double2(P.lut[i], P.lut[512 + i]);
	__syncthreads();

	const double w2 = P.W.w[2];
	long long *corp = P.acc + (uint64_t)ACC_PER_CAND * P.C;
	const uint4 *src = reinterpret_cast<const uint4 *>(P.recs);
	const uint64_t stride = (uint64_t)gridDim.x * K2_THREADS;
	uint64_t t = (uint64_t)blockIdx.x * K2_THREADS + tid;
	// every warp runs the same number of chunk rounds (the flush uses full-warp collectives)
	const uint64_t per_thread = (P.T + stride - 1) / stride;
	const uint64_t rounds = (per_thread + K2_CHUNK - 1) / K2_CHUNK;

	for (uint64_t round = 0; round < rounds; round++) {
		unsigned long long mTot = 0, mGood = 0, mBad = 0;        // 5 modes x 12 bits       (APO:519-525)
		unsigned long long c01a = 0, c01b = 0;                    // (fb + 3*err) x 10 bits, endTime unset / set
		unsigned long long c3 = 0, c4 = 0, c5 = 0, c7 = 0, c8 = 0; // 5 slots x 12 bits, slot 4 = not pushed
		unsigned long long c6 = 0;                                // 7 slots x 9 bits, slot 6 = not pushed
		unsigned long long pat = 0;                               // 6 patterns x 10 bits
		unsigned long long tool[3] = {0, 0, 0};
		long long fxR = 0, fxD2 = 0;
		uint32_t nValid = 0;

		// software pipeline: the next record of this thread is in flight while the current one is scored
		union Rec { uint4 q[2]; apo_record r; } nxt;
		if (t < P.T) { nxt.q[0] = __ldg(src + 2 * t); nxt.q[1] = __ldg(src + 2 * t + 1); }
		for (int it = 0; it < K2_CHUNK && t < P.T; it++, t += stride) {
			Rec u = nxt;
			if (it + 1 < K2_CHUNK && t + stride < P.T) { nxt.q[0] = __ldg(src + 2 * (t + stride)); nxt.q[1] = __ldg(src + 2 * (t + stride) + 1); }
			const apo_record &r = u.r;
			const bool good = r.feedback == 1, bad = r.feedback == 2;
			const uint32_t msh = 12u * (r.mode < APO_NMODE ? r.mode : 0u);               // APO:627-633
			mTot += 1ull << msh; mGood += (unsigned long long)good << msh; mBad += (unsigned long long)bad << msh;
			tool[0] += r.toolCalls; tool[1] += r.toolSucc; tool[2] += r.toolFail;       // TCS:603-605

			if (r.flags & APO_F_VALID) {                                                 // APO:550, TCS:606
				double ws; CatIdx ix;
				const uint32_t mask = record_ws_table_t<true>(r, w2, s_cat, ws, ix);
				const double2 tw = s_lut[mask];
				if (tw.x > 0.0) {                                                        // TCS:784 totalWeight > 0
					fxR += to_fx(div_lut<false>(ws, tw));
					nValid++;
					const unsigned long long one01 = 1ull << (10u * (ix.i01 % 6u));
					if (ix.i01 >= 6u) c01b += one01; else c01a += one01;
					c3 += 1ull << (12u * ix.i3); c4 += 1ull << (12u * ix.i4); c5 += 1ull << (12u * ix.i5);
					c6 += 1ull << (9u * ix.i6); c7 += 1ull << (12u * ix.i7); c8 += 1ull << (12u * ix.i8);
					fxD2 += to_fx(ix.d2);                                                // +0.0 when not pushed
				}
			}
			if (bad) {                                                                   // APO:644-755: every predicate ANDs 'bad'
				const unsigned long long gi = P.idx_base + t;
				const bool hit[APO_NPAT] = {
				    (r.flags & APO_F_ERRORS) != 0,       // P1 APO:644
				    (r.flags & APO_F_FAILSPAN) != 0,     // P2 APO:666-670
				    r.tokens > 10000u,                   // P3 APO:693
				    r.llmCalls > 2u,                     // P4 APO:713
				    r.userMsgs >= 4u,                    // P5 APO:733-734
				    (double)r.toolDurMs > 15000.0,       // P6 APO:754
				};
#pragma unroll
				for (int p = 0; p < APO_NPAT; p++) {
					if (hit[p]) {
						pat += 1ull << (10 * p);
						// slice(0,3): first three in corpus order
						if (gi < *((volatile unsigned long long *)&s_ex[3 * p + 2])) ex_insert(&s_ex[3 * p], gi);
					}
				}
			}
		}

		// ---- flush this chunk: warp-reduce every field, lane 0 turns counts into exact sums
		uint32_t good = 0, badn = 0, total = 0;
#pragma unroll
		for (int m = 0; m < APO_NMODE; m++) {
			const uint32_t a = field_sum<12>(mTot, m), g = field_sum<12>(mGood, m), b = field_sum<12>(mBad, m);
			if (lane == 0) {
				if (a) atomicAdd((unsigned long long *)corp + CORP_MODE + 3 * m, (unsigned long long)a);
				if (g) atomicAdd((unsigned long long *)corp + CORP_MODE + 3 * m + 1, (unsigned long long)g);
				if (b) atomicAdd((unsigned long long *)corp + CORP_MODE + 3 * m + 2, (unsigned long long)b);
			}
			total += a; good += g; badn += b;
		}
		if (lane == 0) {                                                                 // APO:513-516
			if (good) atomicAdd((unsigned long long *)corp + CORP_TALLY, (unsigned long long)good);
			if (badn) atomicAdd((unsigned long long *)corp + CORP_TALLY + 1, (unsigned long long)badn);
			if (total - good - badn) atomicAdd((unsigned long long *)corp + CORP_TALLY + 2, (unsigned long long)(total - good - badn));
		}
#pragma unroll
		for (int p = 0; p < APO_NPAT; p++) {
			const uint32_t n = field_sum<10>(pat, p);
			if (lane == 0 && n) atomicAdd((unsigned long long *)corp + CORP_PAT + p, (unsigned long long)n);
		}
#pragma unroll
		for (int i = 0; i < 3; i++) {
			const unsigned long long s = warp_sum_u64(tool[i]);
			if (lane == 0 && s) atomicAdd((unsigned long long *)corp + CORP_TOOL + i, s);
		}
		{   // finalReward sum and count (APO:550-553)
			Acc128 a; a.lo = (unsigned long long)fxR; a.hi = fxR >> 63;
			flush_acc128(a, corp + CORP_REWARD, lane);
			Acc128 b; b.lo = (unsigned long long)fxD2; b.hi = fxD2 >> 63;
			flush_acc128(b, corp + CORP_DIM + 4 * 2, lane);
		}
		const uint32_t nv = warp_sum_u32(nValid);
		// d0 / d1 from the (feedback, hasErrors, endTime) census: TCS:677-691
		__int128 s0 = 0, s1 = 0;
#pragma unroll
		for (int ended = 0; ended < 2; ended++)
#pragma unroll
			for (int err = 0; err < 2; err++)
#pragma unroll
				for (int fb = 0; fb < 3; fb++) {
					const uint32_t n = field_sum<10>(ended ? c01b : c01a, fb + 3 * err);
					const double d0 = fb == 1 ? 1.0 : (fb == 2 ? -1.0 : 0.0);
					const double d1 = fb == 1 ? 1.0 : (err ? -0.5 : (ended ? 0.8 : 0.5));
					s0 += (__int128)n * to_fx(d0);
					s1 += (__int128)n * to_fx(d1);
				}
		const double lv_rel[4] = {1.0, -0.2, -0.5, -1.0}, lv_cnt[4] = {1.0, 0.3, -0.3, -0.8}, lv_dur[4] = {1.0, 0.5, 0.0, -0.5};
		__int128 s3 = 0, s4 = 0, s5 = 0, s7 = 0, s8 = 0, s6 = 0;
		uint32_t n3 = 0, n5 = 0, n6 = 0, n7 = 0, n8 = 0;
#pragma unroll
		for (int k = 0; k < 4; k++) {
			const uint32_t a3 = field_sum<12>(c3, k), a4 = field_sum<12>(c4, k), a5 = field_sum<12>(c5, k);
			const uint32_t a7 = field_sum<12>(c7, k), a8 = field_sum<12>(c8, k);
			s3 += (__int128)a3 * to_fx(lv_rel[k]); s4 += (__int128)a4 * to_fx(lv_cnt[k]); s5 += (__int128)a5 * to_fx(lv_dur[k]);
			s7 += (__int128)a7 * to_fx(lv_dur[k]); s8 += (__int128)a8 * to_fx(lv_cnt[k]);
			n3 += a3; n5 += a5; n7 += a7; n8 += a8;
		}
#pragma unroll
		for (int k = 0; k < 6; k++) {                                                    // TCS:735: max(-1, 1 - k*0.4)
			const uint32_t a6 = field_sum<9>(c6, k);
			double ef = __dadd_rn(1.0, -__dmul_rn((double)k, 0.4));
			ef = ef < -1.0 ? -1.0 : ef;
			s6 += (__int128)a6 * to_fx(ef);
			n6 += a6;
		}
		if (lane == 0) {
			const __int128 sums[APO_NDIM] = {s0, s1, 0, s3, s4, s5, s6, s7, s8};
			const uint32_t cnts[APO_NDIM] = {nv, nv, n3, n3, n3, n5, n6, n7, n8};            // d2,d3,d4 are pushed together (TCS:695)
#pragma unroll
			for (int i = 0; i < APO_NDIM; i++) {
				if (i != 2) add_i128(corp + CORP_DIM + 4 * i, sums[i]);
				if (cnts[i]) atomicAdd((unsigned long long *)corp + CORP_DIM + 4 * i + 3, (unsigned long long)cnts[i]);
			}
			if (nv) atomicAdd((unsigned long long *)corp + CORP_REWARD + 3, (unsigned long long)nv);
		}
	}
	__syncthreads();
	if (tid < APO_NPAT) {
#pragma unroll
		for (int k = 0; k < 3; k++) {
			const unsigned long long v = s_ex[3 * tid + k];
			if (v != ~0ull && v < P.ex_scratch[3 * tid + 2]) ex_insert(&P.ex_scratch[3 * tid], v);
		}
	}

	// ---- last CTA: publish this rank's examples; at one rank also finalize (fused)
	__threadfence();
	__syncthreads();
	if (tid == 0) {
		const unsigned int ticket = atomicAdd(P.ticket, 1u);
		s_last = (ticket == gridDim.x - 1);
	}
	__syncthreads();
	if (!s_last) return;
	__threadfence();
	if (tid < APO_NPAT * 3) {
		const unsigned long long v = *((volatile unsigned long long *)&P.ex_scratch[tid]);
		corp[CORP_EX + 18 * P.rank + tid] = v == ~0ull ? 0ll : (long long)(v + 1);
	}
	if (tid == 0) { corp[CORP_NREC] = (long long)P.T; *P.ticket = 0; }
	__threadfence();
	__syncthreads();
	if (P.fuse_finalize) finalize_block(P.fin);
}

cudaError_t run_detect6(const K2Params &P, int sm_count, cudaStream_t st) {
	uint64_t want = (P.T + K2_THREADS - 1) / K2_THREADS;
	int grid = sm_count * 2;            // two resident CTAs per SM (launch bounds), one wave
	if ((uint64_t)grid > want) grid = (int)(want ? want : 1);
	k_detect6<<<grid, K2_THREADS, 0, st>>>(P);
	return cudaGetLastError();
}

// =================================================================== K3 finalize: segmented sum + radix top-K
__device__ __forceinline__ unsigned long long score_key(double s) {
	const unsigned long long b = (unsigned long long)__double_as_longlong(s);
	return (b >> 63) ? ~b : (b | 0x8000000000000000ull);     // ascending order-preserving map
}

__device__ void build_report(const FinalizeParams &F, const long long *corp_g) {
	apo_corpus_report &R = *F.report;
	// L2 loads: the partials were produced by other CTAs' atomics (or by the allreduce)
	long long corp[CORP_FIXED];
	for (int i = 0; i < CORP_FIXED; i++) corp[i] = __ldcg(corp_g + i);
	R.total = (uint64_t)corp[CORP_NREC];
	R.good = (uint64_t)corp[CORP_TALLY]; R.bad = (uint64_t)corp[CORP_TALLY + 1]; R.none = (uint64_t)corp[CORP_TALLY + 2];
	const uint64_t twf = R.good + R.bad;
	R.goodRate = twf > 0 ? __ddiv_rn((double)R.good, (double)twf) : 0.0;                     // APO:546-547
	for (int m = 0; m < APO_NMODE; m++) {
		for (int k = 0; k < 3; k++) R.byMode[m][k] = (uint64_t)corp[CORP_MODE + 3 * m + k];
		const uint64_t tot = R.byMode[m][1] + R.byMode[m][2];
		R.byModeGoodRate[m] = tot > 0 ? __ddiv_rn((double)R.byMode[m][1], (double)tot) : 0.0; // APO:541-544
	}
	R.withReward = (uint64_t)corp[CORP_REWARD + 3];
	R.rewardSum = limbs_to_double(corp + CORP_REWARD);
	R.avgReward = R.withReward > 0 ? __ddiv_rn(R.rewardSum, (double)R.withReward) : __longlong_as_double(0x7ff8000000000000ll);
	for (int i = 0; i < APO_NDIM; i++) {
		apo_dimstat &D = R.dim[i];
		D.sum = limbs_to_double(corp + CORP_DIM + 4 * i);
		D.count = (uint64_t)corp[CORP_DIM + 4 * i + 3];
		D.avg = D.count > 0 ? __ddiv_rn(D.sum, (double)D.count) : 0.0;                       // APO:567
		D.low_flag = (D.count >= 5 && D.avg < -0.3) ? 1 : 0;                                 // APO:575
		D.low_severity = D.avg < -0.5 ? 2 : 1;                                               // APO:591
		D.sugg_flag = (D.count >= 3 && D.avg < 0.0) ? 1 : 0;                                 // APO:802
		D.sugg_priority = D.avg < -0.5 ? 2 : 1;                                              // APO:819
		for (int k = 0; k < 4; k++) D.pad[k] = 0;
	}
	const uint64_t minc[APO_NPAT] = {2, 2, 3, 2, 2, 2};                                      // APO:645,671,695,715,736,756
	for (int p = 0; p < APO_NPAT; p++) {
		apo_pattern &Q = R.pat[p];
		Q.count = R.bad == 0 ? 0 : (uint64_t)corp[CORP_PAT + p];                             // APO:641
		Q.flag = Q.count >= minc[p] ? 1 : 0;
		uint8_t sev = 1;
		if (p == 0 || p == 1) sev = Q.count >= 5 ? 2 : 1;                                    // APO:650,676
		else if (p == 3) sev = 2;                                                            // APO:720
		else if (p == 4) sev = Q.count >= 4 ? 2 : 1;                                         // APO:741
		Q.severity = sev;
		for (int k = 0; k < 6; k++) Q.pad[k] = 0;
		// first three in corpus order: ranks hold disjoint ascending index ranges
		int got = 0;
		for (int k = 0; k < 3; k++) Q.examples[k] = -1;
		for (int r = 0; r < F.nranks && got < 3; r++)
			for (int k = 0; k < 3 && got < 3; k++) {
				const long long v = __ldcg(corp_g + CORP_EX + 18 * r + 3 * p + k);
				if (v > 0) Q.examples[got++] = v - 1;
			}
	}
	R.toolCalls = (uint64_t)corp[CORP_TOOL]; R.toolSucc = (uint64_t)corp[CORP_TOOL + 1]; R.toolFail = (uint64_t)corp[CORP_TOOL + 2];
	R.toolSuccessRate = R.toolCalls > 0 ? __ddiv_rn((double)R.toolSucc, (double)R.toolCalls)  // TCS:624
	                                    : __longlong_as_double(0x7ff8000000000000ll);
}

// Block-wide; every thread of the calling block must enter.  Works for any blockDim.x
// that is a multiple of 32 (<= 1024).
__device__ void finalize_block(const FinalizeParams &F) {
	__shared__ unsigned int s_hist[256];
	__shared__ unsigned long long s_prefix;
	__shared__ unsigned int s_need, s_base_gt, s_base_eq, s_warp_gt[32], s_warp_eq[32];
	const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nth >> 5;
	const uint32_t C = F.C;
	const long long *acc = F.acc;

	// 1. segmented sum: limbs -> score[c] = sum / count  (APO:550-553 per candidate)
	for (uint32_t c = tid; c < C; c += nth) {
		long long l[4];
#pragma unroll
		for (int q = 0; q < 4; q++) l[q] = __ldcg(acc + (uint64_t)ACC_PER_CAND * c + q);
		const uint64_t n = (uint64_t)l[3];
		const double s = n > 0 ? __ddiv_rn(limbs_to_double(l), (double)n) : __longlong_as_double(0xfff0000000000000ll);
		F.scores[c] = s;
		F.counts[c] = n;
		F.keys[c] = score_key(s);
	}
	if (tid == 0 && F.with_corpus) build_report(F, acc + (uint64_t)ACC_PER_CAND * C);
	__syncthreads();
	const uint32_t K = F.K < C ? F.K : C;
	if (K == 0) return;

	// 2. radix select (8-bit digits, MSB first): key of the K-th best candidate
	if (tid == 0) { s_prefix = 0; s_need = K; }
	__syncthreads();
	for (int pass = 0; pass < 8; pass++) {
		const int shift = 56 - 8 * pass;
		for (int i = tid; i < 256; i += nth) s_hist[i] = 0;
		__syncthreads();
		const unsigned long long prefix = s_prefix;
		const unsigned long long maskhi = pass == 0 ? 0ull : (~0ull << (shift + 8));
		for (uint32_t c = tid; c < C; c += nth) {
			const unsigned long long k = F.keys[c];
			if ((k & maskhi) == prefix) atomicAdd(&s_hist[(k >> shift) & 255], 1u);
		}
		__syncthreads();
		if (tid == 0) {
			unsigned int need = s_need;
			int d = 255;
			for (; d > 0; d--) {
				if (s_hist[d] >= need) break;
				need -= s_hist[d];
			}
			s_need = need;
			s_prefix = prefix | ((unsigned long long)d << shift);
		}
		__syncthreads();
	}
	const unsigned long long kth = s_prefix;       // K-th largest key
	const unsigned int need_eq = s_need;           // how many candidates equal to it are taken (lowest indices)

	// 3. ordered compaction: all keys > kth, plus the first need_eq keys == kth, by index
	if (tid == 0) { s_base_gt = 0; s_base_eq = 0; }
	__syncthreads();
	const unsigned int n_gt = K - need_eq;
	for (uint32_t c0 = 0; c0 < C; c0 += nth) {
		const uint32_t c = c0 + tid;
		const unsigned long long k = c < C ? F.keys[c] : 0ull;
		const bool gt = c < C && k > kth, eq = c < C && k == kth;
		const unsigned int bgt = __ballot_sync(0xffffffffu, gt), beq = __ballot_sync(0xffffffffu, eq);
		if (lane == 0) { s_warp_gt[warp] = __popc(bgt); s_warp_eq[warp] = __popc(beq); }
		__syncthreads();
		unsigned int ogt = s_base_gt, oeq = s_base_eq;
		for (int w = 0; w < warp; w++) { ogt += s_warp_gt[w]; oeq += s_warp_eq[w]; }
		const unsigned int lt = (1u << lane) - 1u;
		const unsigned int pgt = ogt + __popc(bgt & lt), peq = oeq + __popc(beq & lt);
		if (gt) { F.sel_key[pgt] = k; F.sel_idx[pgt] = (int32_t)c; }
		if (eq && peq < need_eq) { F.sel_key[n_gt + peq] = k; F.sel_idx[n_gt + peq] = (int32_t)c; }
		__syncthreads();
		if (tid == 0) {
			unsigned int a = 0, b = 0;
			for (int w = 0; w < nwarp; w++) { a += s_warp_gt[w]; b += s_warp_eq[w]; }
			s_base_gt += a; s_base_eq += b;
		}
		__syncthreads();
	}

	// 4. order the K winners: score descending, ties -> lower index (rank by counting)
	for (uint32_t i = tid; i < K; i += nth) {
		const unsigned long long ki = F.sel_key[i];
		const int32_t ci = F.sel_idx[i];
		uint32_t rank = 0;
		for (uint32_t j = 0; j < K; j++) {
			const unsigned long long kj = F.sel_key[j];
			rank += (kj > ki || (kj == ki && F.sel_idx[j] < ci)) ? 1u : 0u;
		}
		F.topk[rank] = ci;
	}
}

__global__ void __launch_bounds__(1024) k_finalize(const FinalizeParams F) { finalize_block(F); }

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
