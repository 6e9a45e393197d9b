// apo_abi.cu — the C ABI of libapo_b200.so (include/apo_b200.h): engine state, device
// memory ownership, launch sequencing, the NCCL join of record-axis shards.
//
// One engine = one GPU = one host thread at a time.  No CPU fallback anywhere: every
// compute entry point launches the sm_100a kernels of apo_kernels.cu or fails.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <unistd.h>
#include <new>
#include <algorithm>
#include <string>
#include <utility>
#include <thread>
#include <vector>
#include <vector>

#include "../../include/apo_b200.h"
#include "apo_format.h"
#include "apo_kernels.h"

namespace {

constexpr int ACC_PER_CAND = 4, CORP_FIXED = 68;
inline uint64_t acc_words(uint32_t C, int nranks) { return (uint64_t)ACC_PER_CAND * C + CORP_FIXED + 18ull * nranks; }

std::string g_create_error;

// TCS:766-776 in push order d0..d8
const double kDefaultWeights[APO_NDIM] = {0.25, 0.18, 0.12, 0.08, 0.05, 0.05, 0.08, 0.08, 0.11};

// ---- NCCL through dlopen: the library loads (and every single-GPU path works) without it
struct NcclId { char b[APO_UNIQUE_ID_BYTES]; };
struct NcclApi {
	void *h = nullptr;
	int (*GetUniqueId)(NcclId *) = nullptr;
	int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
	int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
	int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
	int (*CommDestroy)(void *) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool nccl_load(std::string &err) {
	if (g_nccl.h) return true;
	const char *names[] = {"libnccl.so.2", "libnccl.so"};
	for (const char *n : names) {
		g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (g_nccl.h) break;
	}
	if (!g_nccl.h) { err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
	g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.h, "ncclGetUniqueId");
	g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.h, "ncclCommInitRank");
	g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.h, "ncclAllReduce");
	g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(g_nccl.h, "ncclAllGather");
	g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.h, "ncclCommDestroy");
	g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.h, "ncclGetErrorString");
	if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
		err = "libnccl is missing a required symbol"; g_nccl.h = nullptr; return false;
	}
	return true;
}
constexpr int kNcclInt64 = 4, kNcclInt8 = 0, kNcclSum = 0, kNcclMin = 3;
constexpr uint32_t kMaxK = 16384;      // the K winners are ordered by counting (O(K^2)); beam widths are tiny (APO:288: 4)

template <class T>
struct DevBuf {
	T *p = nullptr; uint64_t cap = 0;   // capacity in elements
	cudaError_t reserve(uint64_t n) {
		if (n <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
		if (e == cudaSuccess) cap = n;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

inline uint64_t round_up(uint64_t x, uint64_t m) { return (x + m - 1) / m * m; }

}  // namespace

struct apo_engine {
	int device = 0, sm_count = 148;
	cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
	std::string err;
	apo::Weights W;
	double lut_tw[512], lut_rc[512], lut_cat[apo::CAT_WORDS];
	DevBuf<double> d_lut;          // [0,512) total weight, [512,1024) reciprocal, [1024,1088) categorical products

	DevBuf<apo_record> corpus; uint64_t corpus_T = 0, corpus_base = 0;
	DevBuf<float> dims; const float *dims_ptr = nullptr; uint32_t dims_C = 0; uint64_t dims_T = 0, dims_pitch = 0;
	// Form Q (compact) copy of the evaluations: replaces dims/dims_ptr when `compact` is set
	DevBuf<unsigned long long> q8; DevBuf<float> qd2; DevBuf<unsigned short> qli; DevBuf<uint32_t> qbook;   // qbook: [8][256] codebook + [8] overflow flags
	DevBuf<double> d_ptab, d_pair; DevBuf<float> stage, d_cbf, d_d2book, d_d2stream; uint32_t q_n0 = 0, q_n1 = 0; bool q_pair_ok = false;
	uint32_t qbook_host[8 * 256]; bool compact = false;
	uint32_t d2book_host[4096]; uint32_t d2book_n = ~0u;       // Form P export of the resident tensor (~0 = not built yet)
	DevBuf<uint8_t> roll; uint32_t roll_C = 0, roll_row = 32; uint64_t roll_T = 0, roll_pitch = 0;   // Form R (32 B) or R16 (16 B) rows
	// Form T (csrc/apo_tuple.cu): resident index planes + dictionary; per call: product table of its codebook, table values
	DevBuf<unsigned short> tup_l; DevBuf<unsigned char> tup_h; uint32_t tup_C = 0, tup_n = 0; uint64_t tup_T = 0, tup_pitch = 0;
	DevBuf<uint32_t> tup_pc, tup_bad; DevBuf<unsigned short> tup_pd; DevBuf<float> tup_d2; DevBuf<double> tup_ptab; DevBuf<long long> tup_val;
	uint32_t tup_book[8 * 256]; bool tup_used = false;
	DevBuf<uint32_t> tups_pc; DevBuf<unsigned short> tups_pd; DevBuf<float> tups_d2;   // dictionary of a host-streaming call

	// acc: [acc_words(C, nranks)] partial vector | [18] example scratch (inverted indices) | [1] ticket — one allocation, so one
	// memset arms a scoring call.  acc_joined: the joined vector (> 1 rank); acc keeps this rank's partials.
	DevBuf<long long> acc, acc_joined, acc_snapshot; uint32_t last_C = 0;
	uint64_t acc_clean_words = 0;        // > 0: the first acc_clean_words of acc are known to be zero (left so by the last one-shot call)
	const long long *partials_src = nullptr;   // where apo_debug_partials finds the candidate partials of the last call
	DevBuf<unsigned long long> misc;     // NCCL warm-up / status scratch
	DevBuf<uint8_t> result;              // scores | counts | topk | report
	DevBuf<unsigned long long> keys, sel_key; DevBuf<int32_t> sel_idx;
	uint8_t *h_result = nullptr; uint64_t h_result_cap = 0;
	DevBuf<uint8_t> win[2]; cudaEvent_t win_free[2] = {nullptr, nullptr}, win_ready[2] = {nullptr, nullptr};
	uint8_t *h_stage[2] = {nullptr, nullptr}; uint64_t h_stage_cap = 0; cudaEvent_t stage_done[2] = {nullptr, nullptr};   // pinned staging of pageable callers
	DevBuf<apo_record> batch_in; DevBuf<double> batch_out; DevBuf<uint32_t> batch_mask;

	void *comm = nullptr; int nranks = 1, rank = 0;
	// peer-memory join (csrc/apo_corpus.cuh peer_join): this rank's block and the peer-mapped views of the others
	uint8_t *peer_block = nullptr; uint64_t peer_slot_words = 0; bool peer_ok = false;
	void *peer_base[apo::PEER_MAX] = {nullptr}; bool peer_ipc[apo::PEER_MAX] = {false};
	unsigned long long join_epoch = 0;
	bool env_no_fuse = false, env_force_fuse = false, env_no_staging = false, env_nccl_join = false;   // apo_set_tuning
	double stream_bytes_per_ms = 7.4e9, stream_q_bytes_per_ms = 5.0e9, scan_ms_per_record = 3.35e-7;   // fuse heuristic, derived from the device at apo_create
	size_t mem_pitch_max = 0;
	uint8_t *zc_host = nullptr; uint64_t zc_cap = 0;   // zero-copy block of the single-trace path
	bool timing_on = true;
	cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
	std::vector<cudaEvent_t> k1_ev;      // start/stop pairs of the K1 launches of the current scoring call
	size_t k1_used = 0;
	uint32_t score_C = 0; bool scoring = false;
	apo_timing timing{};
};

namespace {

int fail(apo_engine *e, int code, const char *fmt, ...) {
	char buf[512];
	va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
	if (e) e->err = buf; else g_create_error = buf;
	return code;
}
#define CK(call)                                                                                    \
	do {                                                                                            \
		cudaError_t _c = (call);                                                                    \
		if (_c != cudaSuccess) return fail(e, APO_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_c), __FILE__, __LINE__); \
	} while (0)

void build_luts(apo_engine *e) {
	for (uint32_t m = 0; m < 512; m++) {
		double tw = 0.0;                            // TCS:777-783: totalWeight += w in push order
		for (int i = 0; i < APO_NDIM; i++) if (m & (1u << i)) tw += e->W.w[i];
		if (!(tw > 0.0)) tw = -1.0;                 // totalWeight == 0: finalReward stays null (TCS:784); a negative entry tells the kernels not to count it
		e->lut_tw[m] = tw;
		// K1 divides by multiplying with RN(1/tw) + one FMA correction (Markstein); that is only
		// proven correctly rounded when the significand of tw is not all ones — flag those masks
		// with a negative reciprocal so the kernel takes the generic IEEE division for them.
		uint64_t bits; memcpy(&bits, &tw, 8);
		const bool all_ones = (bits & 0xFFFFFFFFFFFFFull) == 0xFFFFFFFFFFFFFull;
		e->lut_rc[m] = tw < 0.0 ? 1.0 : (all_ones ? -(1.0 / tw) : 1.0 / tw);
	}
	// categorical product table (csrc/apo_device.cuh CAT_*): value * weight with the same IEEE
	// multiplications the reference performs at TCS:781, tabulated once per weight vector
	const double *w = e->W.w;
	double *c = e->lut_cat;
	for (int i = 0; i < apo::CAT_WORDS; i++) c[i] = 0.0;
	for (int ended = 0; ended < 2; ended++)
		for (int err = 0; err < 2; err++)
			for (int fb = 0; fb < 3; fb++) {
				const double d0 = fb == 1 ? 1.0 : (fb == 2 ? -1.0 : 0.0);                       // TCS:677-678
				double d1 = 0.5;                                                                 // TCS:682-691
				if (ended && !err) d1 = 0.8;
				if (err) d1 = -0.5;
				if (fb == 1) d1 = 1.0;
				volatile double s = 0.0;
				s = s + d0 * w[0];
				s = s + d1 * w[1];
				c[0 + fb + 3 * err + 6 * ended] = s;
			}
	const double lv_rel[4] = {1.0, -0.2, -0.5, -1.0}, lv_cnt[4] = {1.0, 0.3, -0.3, -0.8}, lv_dur[4] = {1.0, 0.5, 0.0, -0.5};
	for (int k = 0; k < 4; k++) {
		c[12 + k] = lv_rel[k] * w[3]; c[17 + k] = lv_cnt[k] * w[4]; c[22 + k] = lv_dur[k] * w[5];
		c[34 + k] = lv_dur[k] * w[7]; c[39 + k] = lv_cnt[k] * w[8];
	}
	for (int k = 0; k < 6; k++) {
		volatile double over = (double)k;
		volatile double m = over * 0.4;
		double ef = 1 - m;                                                                       // TCS:735
		if (ef < -1) ef = -1;
		c[27 + k] = ef * w[6];
	}
	// direct tables of K1r: [agent][min(counter, cap)] -> the level product above (bit copies)
	using namespace apo;
	for (int ag = 0; ag < 2; ag++) {
		const uint32_t sev = ag ? 5u : 3u, mod = ag ? 3u : 2u, mnr = ag ? 2u : 1u;                       // TCS:702-704
		for (uint32_t f = 0; f <= 5; f++) c[DIR_D3 + ag * 7 + f] = c[CAT_D3 + (f >= mnr) + (f >= mod) + (f >= sev)];
		c[DIR_D3 + ag * 7 + 6] = 0.0;
		const uint32_t cexc = ag ? 8u : 3u, cgood = ag ? 15u : 6u, cfair = ag ? 25u : 10u;               // TCS:711-713
		c[DIR_D4 + ag * 27] = 0.0;
		for (uint32_t n = 1; n <= 26; n++) c[DIR_D4 + ag * 27 + n] = c[CAT_D4 + (n > cexc) + (n > cgood) + (n > cfair)];
		const uint32_t thr6 = ag ? 3u : 1u;                                                               // TCS:734
		c[DIR_D6 + ag * 10] = 0.0;
		for (uint32_t n = 1; n <= 9; n++) { const uint32_t over = n > thr6 ? n - thr6 : 0u; c[DIR_D6 + ag * 10 + n] = c[CAT_D6 + (over < 5u ? over : 5u)]; }
		const uint32_t thr8 = ag ? 3u : 2u;                                                               // TCS:756
		c[DIR_D8 + ag * 11] = 0.0;
		for (uint32_t n = 1; n <= 10; n++) c[DIR_D8 + ag * 11 + n] = c[CAT_D8 + (n > thr8) + (n > 2 * thr8) + (n > 3 * thr8)];
	}
}

int upload_luts(apo_engine *e) {
	CK(e->d_lut.reserve(1024 + apo::CAT_WORDS));
	CK(cudaMemcpyAsync(e->d_lut.p + 1024, e->lut_cat, apo::CAT_WORDS * 8, cudaMemcpyHostToDevice, e->stream));
	CK(cudaMemcpyAsync(e->d_lut.p, e->lut_tw, 512 * 8, cudaMemcpyHostToDevice, e->stream));
	CK(cudaMemcpyAsync(e->d_lut.p + 512, e->lut_rc, 512 * 8, cudaMemcpyHostToDevice, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return APO_OK;
}

// value*weight per code for the eight coded dimensions (csrc/apo_compact.cu); table 0 = fl(0 + d0*w0)
std::vector<double> make_ptab(const uint32_t *book, const apo::Weights &W) {
	static const int dim_of[8] = {0, 1, 3, 4, 5, 6, 7, 8};
	std::vector<double> tab(8 * 256, 0.0);
	for (int j = 0; j < 8; j++)
		for (int c = 0; c < 255; c++) {
			const uint32_t bits = book[256 * j + c];
			if (bits == 0xFFFFFFFFu) continue;
			float f; memcpy(&f, &bits, 4);
			volatile double p = (double)f * W.w[dim_of[j]];             // TCS:781 value * weight
			if (j == 0) { volatile double z = 0.0; p = z + p; }          // weightedSum = 0 + first product
			tab[256 * j + c] = p;
		}
	return tab;
}

int upload_ptab(apo_engine *e) {
	const std::vector<double> tab = make_ptab(e->qbook_host, e->W);
	CK(e->d_ptab.reserve(8 * 256));
	CK(cudaMemcpyAsync(e->d_ptab.p, tab.data(), 8 * 256 * 8, cudaMemcpyHostToDevice, e->stream));
	// tables of the mixed-lookup variants: fp32 values per code, and the exact two-term prefix of dimensions 0 and 1
	std::vector<float> cbf(8 * 256, 0.0f);
	uint32_t used[8] = {0};
	for (int j = 0; j < 8; j++)
		for (int c = 0; c < 255; c++) {
			const uint32_t bits = e->qbook_host[256 * j + c];
			if (bits == 0xFFFFFFFFu) continue;
			memcpy(&cbf[256 * j + c], &bits, 4);
			used[j] = (uint32_t)c + 1;                                // dense codes: 0 .. used-1
		}
	e->q_n0 = used[0]; e->q_n1 = used[1];
	e->q_pair_ok = (uint64_t)(used[0] + 1) * (used[1] + 1) <= (uint64_t)apo::KQ_PAIR_MAX;
	std::vector<double> pair(apo::KQ_PAIR_MAX, 0.0);
	if (e->q_pair_ok)
		for (uint32_t a = 0; a <= used[0]; a++)
			for (uint32_t b = 0; b <= used[1]; b++) {
				volatile double s = tab[a < used[0] ? a : 255];           // fl(0 + d0*w0), +0.0 when absent
				s = s + tab[256 + (b < used[1] ? b : 255)];               // + d1*w1 (TCS:781, second push)
				pair[a * (used[1] + 1) + b] = s;
			}
	CK(e->d_cbf.reserve(8 * 256));
	CK(e->d_pair.reserve(apo::KQ_PAIR_MAX));
	CK(cudaMemcpyAsync(e->d_cbf.p, cbf.data(), 8 * 256 * 4, cudaMemcpyHostToDevice, e->stream));
	CK(cudaMemcpyAsync(e->d_pair.p, pair.data(), apo::KQ_PAIR_MAX * 8, cudaMemcpyHostToDevice, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return APO_OK;
}

int compact_begin(apo_engine *e, uint32_t C, uint64_t pitch) {
	CK(e->q8.reserve((uint64_t)(C ? C : 1) * pitch));
	CK(e->qd2.reserve((uint64_t)(C ? C : 1) * pitch));
	CK(e->qli.reserve((uint64_t)(C ? C : 1) * pitch));
	CK(e->qbook.reserve(8 * 256 + 8));
	CK(cudaMemsetAsync(e->qbook.p, 0xFF, 8 * 256 * 4, e->stream));
	CK(cudaMemsetAsync(e->qbook.p + 8 * 256, 0, 8 * 4, e->stream));
	// pad evaluations of every row: all codes 255 + NaN = finalReward null
	CK(cudaMemsetAsync(e->q8.p, 0xFF, (uint64_t)(C ? C : 1) * pitch * 8, e->stream));
	CK(cudaMemsetAsync(e->qd2.p, 0xFF, (uint64_t)(C ? C : 1) * pitch * 4, e->stream));
	return APO_OK;
}

// returns APO_E_STATE when a coded dimension has more than 255 distinct values
int compact_finish(apo_engine *e, uint32_t C, uint64_t T, uint64_t pitch) {
	uint32_t host[8 * 256 + 8];
	CK(cudaMemcpyAsync(host, e->qbook.p, sizeof host, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	for (int j = 0; j < 8; j++)
		if (host[8 * 256 + j]) {
			e->q8.release(); e->qd2.release(); e->qli.release();
			return fail(e, APO_E_STATE, "evaluations are not categorical: coded dimension %d has more than 255 distinct values; Form D is kept", j < 2 ? j : j + 1);
		}
	// hash-slot codes -> dense codes ordered by value (deterministic, bank-conflict-free table reads)
	uint8_t remap[8 * 256];
	uint32_t dense[8 * 256];
	for (int j = 0; j < 8; j++) {
		std::vector<std::pair<float, int>> used;
		for (int s = 0; s < 255; s++) {
			const uint32_t bits = host[256 * j + s];
			if (bits != 0xFFFFFFFFu) { float f; memcpy(&f, &bits, 4); used.push_back({f, s}); }
		}
		std::sort(used.begin(), used.end(), [&](const std::pair<float, int> &a, const std::pair<float, int> &b) {
			if (a.first != b.first) return a.first < b.first;
			return host[256 * j + a.second] > host[256 * j + b.second];      // -0.0 (sign bit set) before +0.0
		});
		for (int c = 0; c < 256; c++) { remap[256 * j + c] = 255; dense[256 * j + c] = 0xFFFFFFFFu; }
		for (size_t r = 0; r < used.size(); r++) { remap[256 * j + used[r].second] = (uint8_t)r; dense[256 * j + r] = host[256 * j + used[r].second]; }
	}
	CK(e->stage.reserve(1024));
	CK(cudaMemcpyAsync(e->stage.p, remap, sizeof remap, cudaMemcpyHostToDevice, e->stream));
	CK(apo::run_recode(e->q8.p, e->qd2.p, e->qli.p, (uint64_t)C * pitch, (const uint8_t *)e->stage.p, e->stream));
	CK(cudaMemcpyAsync(e->qbook.p, dense, sizeof dense, cudaMemcpyHostToDevice, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	memcpy(e->qbook_host, dense, sizeof e->qbook_host);
	int rc = upload_ptab(e);
	if (rc) return rc;
	e->compact = true; e->dims_C = C; e->dims_T = T; e->dims_pitch = pitch; e->dims_ptr = nullptr;
	e->d2book_n = ~0u;
	return APO_OK;
}

// true when the driver does not know the pointer: ordinary pageable host memory (malloc, JS ArrayBuffer, numpy)
bool is_pageable(const void *p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
	return a.type == cudaMemoryTypeUnregistered;
}

// rows x width bytes, strided -> strided, split over a few host threads (a single memcpy stream does not fill PCIe 5)
void parallel_rows_copy(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch, size_t width, uint32_t rows) {
	unsigned nt = std::thread::hardware_concurrency();
	static const int env_threads = [] { const char *v = getenv("APO_HOST_THREADS"); return v ? atoi(v) : 0; }();   // read once per process
	if (env_threads > 0) nt = (unsigned)env_threads;
	else nt = nt ? (nt > 16 ? 16 : nt) : 4;        // default: at most 16 (measured: 4 / 8 / 16 / 32 threads -> 15 / 22-29 / 30 / 25 GB/s); APO_HOST_THREADS overrides
	if (nt > 64) nt = 64;
	const size_t total = width * rows;
	if (nt == 1 || total < (8u << 20)) { for (uint32_t r = 0; r < rows; r++) memcpy(dst + r * dpitch, src + r * spitch, width); return; }
	// split every row into nt column slices: works for C = 1 as well as for many short rows
	auto work = [&](unsigned k) {
		const size_t lo = width * k / nt, hi = width * (k + 1) / nt;
		for (uint32_t r = 0; r < rows; r++) memcpy(dst + r * dpitch + lo, src + r * spitch + lo, hi - lo);
	};
	std::vector<std::thread> th;
	unsigned started = 1;
	try {
		for (unsigned k = 1; k < nt; k++, started++) th.emplace_back(work, k);
	} catch (...) {}                                   // thread limit reached: the slices not handed out are copied here
	work(0);
	for (unsigned k = started; k < nt; k++) work(k);
	for (auto &t : th) t.join();
}

// two page-locked staging buffers of at least `bytes` each (kept for the life of the handle)
int ensure_stage(apo_engine *e, uint64_t bytes) {
	if (e->h_stage_cap >= bytes) return APO_OK;
	for (int i = 0; i < 2; i++) { if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]); e->h_stage[i] = nullptr; }
	e->h_stage_cap = 0;
	for (int i = 0; i < 2; i++) CK(cudaMallocHost((void **)&e->h_stage[i], bytes));
	e->h_stage_cap = bytes;
	return APO_OK;
}

// cudaMemcpy2DAsync rejects pitches above cudaDeviceProp::memPitch (2 GiB - 1 on B200): rows of a host tensor whose
// record axis is longer than that (T * row bytes > 2 GiB, the "larger than device memory" case of the streaming calls)
// are copied one by one instead.
int copy_rows_h2d(apo_engine *e, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, uint32_t rows, cudaStream_t st) {
	if (rows == 0 || width == 0) return APO_OK;
	if (rows == 1) { CK(cudaMemcpyAsync(dst, src, width, cudaMemcpyHostToDevice, st)); return APO_OK; }
	if (e->mem_pitch_max == 0 || (dpitch <= e->mem_pitch_max && spitch <= e->mem_pitch_max)) {
		CK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, st));
		return APO_OK;
	}
	for (uint32_t r = 0; r < rows; r++)
		CK(cudaMemcpyAsync((uint8_t *)dst + (size_t)r * dpitch, (const uint8_t *)src + (size_t)r * spitch, width, cudaMemcpyHostToDevice, st));
	return APO_OK;
}

// rows x width bytes host -> device on `st` (asynchronous for page-locked sources; a pageable source is fully consumed
// when this returns).  Large pageable sources go through the staging buffers in column slices of all rows.
int h2d_rows(apo_engine *e, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, uint32_t rows, cudaStream_t st) {
	const uint64_t total = (uint64_t)width * rows;
	if (total < (64ull << 20) || !is_pageable(src) || e->env_no_staging)
		return copy_rows_h2d(e, dst, dpitch, src, spitch, width, rows, st);
	const uint64_t cap = 256ull << 20;
	int rc = ensure_stage(e, cap);
	if (rc) return rc;
	size_t W = (size_t)(cap / rows) & ~(size_t)255;            // slice width per row, 256-byte granular
	if (W == 0) W = 256;
	int k = 0;
	for (size_t w0 = 0; w0 < width; w0 += W, k++) {
		const int b = k & 1;
		const size_t w = width - w0 < W ? width - w0 : W;
		if (k >= 2) CK(cudaEventSynchronize(e->stage_done[b]));
		parallel_rows_copy(e->h_stage[b], W, (const uint8_t *)src + w0, spitch, w, rows);
		{ const int rc2 = copy_rows_h2d(e, (uint8_t *)dst + w0, dpitch, e->h_stage[b], W, w, rows, st); if (rc2) return rc2; }
		CK(cudaEventRecord(e->stage_done[b], st));
	}
	return APO_OK;
}

struct ResultLayout { uint64_t off_scores, off_counts, off_topk, off_report, off_meta, bytes; };
ResultLayout result_layout(uint32_t C, uint32_t K) {
	ResultLayout L;
	L.off_scores = 0;
	L.off_counts = L.off_scores + 8ull * C;
	L.off_topk = L.off_counts + 8ull * C;
	L.off_report = round_up(L.off_topk + 4ull * K, 16);
	L.off_meta = round_up(L.off_report + sizeof(apo_corpus_report), 16);
	L.bytes = L.off_meta + sizeof(apo::ResultMeta);
	return L;
}
static_assert(sizeof(apo::ResultMeta) == 32, "result block stays a multiple of 16 bytes");
constexpr uint64_t kHostOutMax = 64ull << 10;      // result blocks up to this size are written to the caller's page-locked buffer by the kernel itself
constexpr uint64_t kTimingMinBytes = 64ull << 20;  // per-stage CUDA events are recorded for calls streaming at least this much (or on request)
constexpr uint64_t ARM_WORDS = 19;                  // 18 example slots + ticket, right behind the partial vector

int ensure_scratch(apo_engine *e, uint32_t C, uint32_t K) {
	if (acc_words(C, e->nranks) + ARM_WORDS > e->acc.cap) e->acc_clean_words = 0;      // a fresh allocation is not zeroed
	CK(e->acc.reserve(acc_words(C, e->nranks) + ARM_WORDS));
	CK(e->acc_snapshot.reserve((uint64_t)ACC_PER_CAND * (C ? C : 1)));
	const ResultLayout L = result_layout(C, K);
	CK(e->result.reserve(L.bytes));
	CK(e->keys.reserve(C ? C : 1));
	CK(e->sel_key.reserve(K ? K : 1));
	CK(e->sel_idx.reserve(K ? K : 1));
	if (e->nranks > 1) CK(e->acc_joined.reserve(acc_words(C, e->nranks)));
	if (e->h_result_cap < L.bytes) {
		if (e->h_result) cudaFreeHost(e->h_result);
		e->h_result = nullptr; e->h_result_cap = 0;
		CK(cudaMallocHost((void **)&e->h_result, L.bytes));
		e->h_result_cap = L.bytes;
	}
	return APO_OK;
}

// the peer-memory join is used when the block exists on every rank and the partial vector fits its slots
bool peer_join_active(const apo_engine *e, uint32_t C) {
	return e->nranks > 1 && e->peer_ok && !e->env_nccl_join && acc_words(C, e->nranks) <= e->peer_slot_words;
}

// layout of one rank's peer block: [PEER_MAX] arrival flags | [PEER_MAX] done-reading flags (128 B) | slot 0 | slot 1
constexpr uint64_t kPeerHeader = 2 * apo::PEER_MAX * 8;
inline unsigned long long *peer_flags(void *base) { return (unsigned long long *)base; }
inline long long *peer_slot(void *base, uint64_t slot_words, int parity) { return (long long *)((uint8_t *)base + kPeerHeader) + (uint64_t)parity * slot_words; }

apo::FinalizeParams make_fin(apo_engine *e, uint32_t C, uint32_t K, int with_corpus) {
	const ResultLayout L = result_layout(C, K);
	apo::FinalizeParams F{};
	F.acc = e->acc.p; F.C = C; F.K = K; F.nranks = e->nranks; F.with_corpus = with_corpus;
	F.scores = (double *)(e->result.p + L.off_scores);
	F.counts = (uint64_t *)(e->result.p + L.off_counts);
	F.keys = e->keys.p; F.sel_key = e->sel_key.p; F.sel_idx = e->sel_idx.p;
	F.topk = (int32_t *)(e->result.p + L.off_topk);
	F.report = (apo_corpus_report *)(e->result.p + L.off_report);
	F.result_base = e->result.p; F.result_bytes = (uint32_t)L.bytes; F.meta_off = (uint32_t)L.off_meta;
	F.host_out = L.bytes <= kHostOutMax ? e->h_result : nullptr;
	F.join.nranks = 1;
	if (peer_join_active(e, C)) {
		apo::JoinParams &J = F.join;
		J.nranks = e->nranks; J.rank = e->rank; J.words = (uint32_t)acc_words(C, e->nranks); J.epoch = e->join_epoch;
		for (int r = 0; r < e->nranks; r++) {
			J.slot[r] = peer_slot(e->peer_base[r], e->peer_slot_words, (int)(e->join_epoch & 1ull));
			J.flag[r] = peer_flags(e->peer_base[r]);
		}
		J.joined = e->acc_joined.p;
	}
	return F;
}

// zero the candidate accumulators of a scoring session
int begin_score(apo_engine *e, uint32_t C, bool whole = false) {
	// whole: also the corpus block, the example scratch and the ticket (one-shot calls: a single memset arms everything —
	// and none at all when the previous one-shot call left the block zeroed, see FinalizeParams::clean_ptr)
	const uint64_t words = whole ? acc_words(C, e->nranks) + ARM_WORDS : (uint64_t)ACC_PER_CAND * C;
	if (!(whole && e->acc_clean_words >= words)) CK(cudaMemsetAsync(e->acc.p, 0, words * 8, e->stream));
	e->acc_clean_words = 0;                         // from here on the block is in use
	e->partials_src = e->nranks > 1 ? e->acc_joined.p : e->acc.p;
	e->last_C = C; e->score_C = C; e->scoring = true; e->k1_used = 0;
	e->timing = apo_timing{};
	if (e->timing_on) CK(cudaEventRecord(e->ev[0], e->stream));
	return APO_OK;
}

int record_k1_event(apo_engine *e, int which) {
	if (!e->timing_on) return APO_OK;
	if (which == 0 && e->k1_used + 2 > e->k1_ev.size()) {
		for (int i = 0; i < 2; i++) { cudaEvent_t ev; CK(cudaEventCreate(&ev)); e->k1_ev.push_back(ev); }
	}
	CK(cudaEventRecord(e->k1_ev[e->k1_used + which], e->stream));
	if (which == 1) e->k1_used += 2;
	return APO_OK;
}

// Form T: finalReward of every dictionary entry for the current weights (device work on e->stream, no synchronisation)
int tuples_values(apo_engine *e, const uint32_t *book_host, const uint32_t *d_pc, const unsigned short *d_pd, uint32_t n, const float *d_d2, bool recip) {
	const std::vector<double> tab = make_ptab(book_host, e->W);
	CK(e->tup_ptab.reserve(8 * 256));
	CK(e->tup_val.reserve((uint64_t)n + 1));
	CK(e->tup_bad.reserve(1));
	CK(cudaMemcpyAsync(e->tup_ptab.p, tab.data(), 8 * 256 * 8, cudaMemcpyHostToDevice, e->stream));   // pageable source: consumed when this returns
	if (!e->tup_used) CK(cudaMemsetAsync(e->tup_bad.p, 0, 4, e->stream));       // a session keeps the flags of its earlier launches
	apo::TupleValParams V{};
	V.tb_pc = d_pc; V.tb_pd = d_pd; V.n = n; V.ptab = e->tup_ptab.p; V.lut = e->d_lut.p; V.d2book = d_d2; V.w2 = e->W.w[2];
	V.tval = e->tup_val.p; V.bad = e->tup_bad.p;
	CK(apo::run_tuple_values(V, recip, e->stream));
	e->timing.launches++;
	e->tup_used = true;
	return APO_OK;
}

// after the scoring call has been synchronised: did K1t meet an index outside the dictionary, or an entry it cannot hold?
int tuples_check(apo_engine *e) {
	e->tup_used = false;
	uint32_t bad = 0;
	CK(cudaMemcpyAsync(&bad, e->tup_bad.p, 4, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	if (bad & 2u) return fail(e, APO_E_ARG, "a dictionary entry is not representable in Form T (finalReward outside (-2, 2), or a tool_success_rate index without a value)");
	if (bad & 1u) return fail(e, APO_E_ARG, "a tuple index is outside the dictionary");
	return APO_OK;
}

// one K1 launch over [first, first+count) of the loaded source into candidates [cand_offset, ...)
int launch_k1_resident(apo_engine *e, const apo_score_opts *o, uint32_t cand_offset, uint64_t first, uint64_t count,
                       const apo::K2Params *fused = nullptr) {
	const bool raw = o->source == APO_SRC_ROLLOUTS;
	const uint32_t C = raw ? e->roll_C : e->dims_C;
	int rc;
	if (o->source == APO_SRC_TUPLES) {
		if (!count) return APO_OK;
		if (first % 8) return fail(e, APO_E_ARG, "window start must be a multiple of 8 for Form T");
		if ((rc = tuples_values(e, e->tup_book, e->tup_pc.p, e->tup_pd.p, e->tup_n, e->tup_d2.p, (o->flags & APO_SCORE_RECIP) != 0))) return rc;
		apo::KtParams P{};
		P.tl = e->tup_l.p + first; P.th = e->tup_h.p + first; P.pitch_evals = e->tup_pitch; P.C = e->tup_C; P.T = count;
		P.acc = e->acc.p + (uint64_t)ACC_PER_CAND * cand_offset; P.tval = e->tup_val.p; P.n_tuples = e->tup_n; P.bad = e->tup_bad.p;
		if ((rc = record_k1_event(e, 0))) return rc;
		CK(apo::run_reward9t(P, e->sm_count, e->stream));
		if ((rc = record_k1_event(e, 1))) return rc;
		e->timing.launches++;
		return APO_OK;
	}
	if (!raw && e->compact) {
		if (!count) return APO_OK;
		if (first % 8) return fail(e, APO_E_ARG, "window start must be a multiple of 8 for the compact layout");
		apo::KqParams Q{};
		Q.q8 = e->q8.p + first; Q.d2 = e->qd2.p + first; Q.li = e->qli.p + first; Q.pitch_evals = e->dims_pitch; Q.C = C; Q.T = count;
		Q.acc = e->acc.p + (uint64_t)ACC_PER_CAND * cand_offset; Q.lut = e->d_lut.p; Q.ptab = e->d_ptab.p; Q.w2 = e->W.w[2];
		Q.pair = e->d_pair.p; Q.cbf = e->d_cbf.p; Q.n0 = e->q_n0; Q.n1 = e->q_n1;
		{
			static const int dim_of[8] = {0, 1, 3, 4, 5, 6, 7, 8};
			for (int j = 0; j < 8; j++) Q.wq[j] = e->W.w[dim_of[j]];
		}
		int qvariant = (int)o->variant;
		if ((qvariant == 0 || qvariant >= 5) && !e->q_pair_ok) qvariant = 4;   // prefix table would not fit: product tables only
		if (fused) { Q.corpus_on = 1; Q.corpus = *fused; }
		if ((rc = record_k1_event(e, 0))) return rc;
		CK(apo::run_reward9q(Q, qvariant, (o->flags & APO_SCORE_RECIP) != 0, e->sm_count, e->stream));
		if ((rc = record_k1_event(e, 1))) return rc;
		e->timing.launches++;
		return APO_OK;
	}
	apo::K1Params P{};
	const int row = raw ? (int)e->roll_row : 36;
	const uint64_t pitch = raw ? e->roll_pitch : e->dims_pitch;
	P.base = (raw ? (const uint8_t *)e->roll.p : (const uint8_t *)e->dims_ptr) + first * row;
	P.pitch_bytes = pitch * row;
	P.C = C; P.T = count; P.acc = e->acc.p + (uint64_t)ACC_PER_CAND * cand_offset;
	P.lut = e->d_lut.p;
	P.W = e->W;
	if (fused) { P.corpus_on = 1; P.corpus = *fused; }
	if (!count) return APO_OK;
	if ((rc = record_k1_event(e, 0))) return rc;
	CK(apo::run_reward9(P, row, (int)o->variant, (o->flags & APO_SCORE_RECIP) != 0, e->sm_count, e->stream));
	if ((rc = record_k1_event(e, 1))) return rc;
	e->timing.launches++;
	return APO_OK;
}

// K2 (+fused finalize) / allreduce / K3, then bring the result block home
bool wants_corpus(const apo_engine *e, const apo_score_opts *o) { return (o->flags & APO_SCORE_CORPUS) && e->corpus_T > 0; }

// the corpus block, the example scratch and the ticket belong to one corpus pass only, so that a scoring session can be
// finished repeatedly while evaluations keep being accumulated (incremental scoring); they sit behind the candidate
// accumulators in one allocation: one memset
int arm_corpus(apo_engine *e, uint32_t C) {
	CK(cudaMemsetAsync(e->acc.p + (uint64_t)ACC_PER_CAND * C, 0, (CORP_FIXED + 18ull * e->nranks + ARM_WORDS) * 8, e->stream));
	return APO_OK;
}

apo::K2Params make_k2(apo_engine *e, uint32_t C, const apo::FinalizeParams &F) {
	apo::K2Params P{};
	unsigned long long *arm = (unsigned long long *)(e->acc.p + acc_words(C, e->nranks));
	P.recs = e->corpus.p; P.T = e->corpus_T; P.idx_base = e->corpus_base; P.C = C; P.rank = e->rank;
	P.acc = e->acc.p; P.ex_scratch = arm; P.ticket = (unsigned int *)(arm + 18);
	P.lut = e->d_lut.p; P.W = e->W; P.fuse_finalize = (e->nranks == 1 || F.join.nranks > 1) ? 1 : 0; P.fin = F;
	return P;
}

// fused_corpus: the scoring kernel already ran the corpus scan and the finalisation
int finish_score(apo_engine *e, const apo_score_opts *o, uint32_t C, double *scores, uint64_t *counts, int32_t *topk,
                 apo_corpus_report *report, bool fused_corpus = false) {
	const uint32_t K = o->K;
	const bool with_corpus = wants_corpus(e, o);
	const bool tm = e->timing_on;
	apo::FinalizeParams F = make_fin(e, C, K, with_corpus ? 1 : 0);
	const bool peer = F.join.nranks > 1;
	if (!fused_corpus) { int rc = arm_corpus(e, C); if (rc) return rc; }
	if (tm) CK(cudaEventRecord(e->ev[1], e->stream));
	bool finalized = fused_corpus && (e->nranks == 1 || peer);
	if (with_corpus && !fused_corpus) {
		const apo::K2Params P = make_k2(e, C, F);
		CK(apo::run_detect6(P, e->sm_count, e->stream));
		e->timing.launches++;
		finalized = P.fuse_finalize != 0;
	}
	if (tm) CK(cudaEventRecord(e->ev[2], e->stream));
	if (e->nranks > 1 && !peer) {
		const int rc = g_nccl.AllReduce(e->acc.p, e->acc_joined.p, (size_t)acc_words(C, e->nranks), kNcclInt64, kNcclSum, e->comm, e->stream);
		if (rc != 0) return fail(e, APO_E_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
		e->timing.launches++;
		F.acc = e->acc_joined.p;
	}
	if (tm) CK(cudaEventRecord(e->ev[3], e->stream));
	if (!finalized) {
		CK(apo::run_finalize(F, e->stream));
		e->timing.launches++;
	}
	if (tm) CK(cudaEventRecord(e->ev[4], e->stream));
	const ResultLayout L = result_layout(C, K);
	if (!F.host_out) CK(cudaMemcpyAsync(e->h_result, e->result.p, L.bytes, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	apo::ResultMeta meta;
	memcpy(&meta, e->h_result + L.off_meta, sizeof meta);
	if (meta.status != 0)
		return fail(e, APO_E_NCCL, "peer-memory join timed out after %.0f ms: a rank did not reach the join of call %llu", meta.join_wait_us * 1e-3,
		            (unsigned long long)e->join_epoch);
	if (scores) memcpy(scores, e->h_result + L.off_scores, 8ull * C);
	if (counts) memcpy(counts, e->h_result + L.off_counts, 8ull * C);
	if (topk && K) memcpy(topk, e->h_result + L.off_topk, 4ull * (K < C ? K : C));
	if (report) {
		if (with_corpus) memcpy(report, e->h_result + L.off_report, sizeof(apo_corpus_report));
		else memset(report, 0, sizeof(apo_corpus_report));
	}
	e->timing.join_wait_ms = meta.join_wait_us * 1e-3f;
	e->timing.join_reduce_ms = meta.join_reduce_us * 1e-3f;
	e->timing.tail_finalize_ms = meta.finalize_us * 1e-3f;
	e->timing.tail_publish_ms = meta.publish_us * 1e-3f;
	if (tm) {
		float ms = 0;
		e->timing.reward_ms = 0;
		for (size_t i = 0; i + 1 < e->k1_used; i += 2) { cudaEventElapsedTime(&ms, e->k1_ev[i], e->k1_ev[i + 1]); e->timing.reward_ms += ms; }
		cudaEventElapsedTime(&ms, e->ev[1], e->ev[2]); e->timing.corpus_ms = ms;
		cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]); e->timing.allreduce_ms = ms;
		cudaEventElapsedTime(&ms, e->ev[3], e->ev[4]); e->timing.finalize_ms = ms;
		cudaEventElapsedTime(&ms, e->ev[0], e->ev[4]); e->timing.total_ms = ms;
	}
	if (e->tup_used) return tuples_check(e);
	return APO_OK;
}

// per-stage CUDA events cost ~1 us each on the host: record them for calls that stream enough for it not to matter, or on request
void choose_timing(apo_engine *e, const apo_score_opts *o, uint64_t bytes) {
	e->timing_on = (o->flags & APO_SCORE_TIMING) != 0 || bytes >= kTimingMinBytes;
}

int check_opts(apo_engine *e, const apo_score_opts *o, uint32_t C, uint64_t T, uint64_t *first, uint64_t *count) {
	if (!o) return fail(e, APO_E_ARG, "opts is NULL");
	if (o->K > C) return fail(e, APO_E_ARG, "K=%u exceeds the number of candidates C=%u", o->K, C);
	if (o->K > kMaxK) return fail(e, APO_E_ARG, "K=%u exceeds the supported beam width %u", o->K, kMaxK);
	if (o->first % 4) return fail(e, APO_E_ARG, "window start must be a multiple of 4");
	if (o->first > T || (o->count && o->first + o->count > T)) return fail(e, APO_E_ARG, "window outside [0,%llu)", (unsigned long long)T);
	*first = o->first;
	*count = o->count ? o->count : T - o->first;
	return APO_OK;
}

}  // namespace

namespace {
struct PeerInfo {                        // exchanged with ncclAllGather at apo_comm_init, 128 bytes per rank
	cudaIpcMemHandle_t handle;           // 64 B
	uint64_t pid, ptr, host_hash;
	int32_t device, ok;
	uint8_t pad[128 - 64 - 24 - 8];
};
static_assert(sizeof(PeerInfo) == 128, "PeerInfo is exchanged as 128 raw bytes");

void peer_teardown(apo_engine *e) {
	if (e->peer_ok && e->peer_block && e->join_epoch > 0) {
		// peers read this block with NVLink loads inside their own launches: wait (bounded) until every peer has signalled
		// that it finished reading the last joined call before the memory goes away
		unsigned long long done[apo::PEER_MAX];
		for (int spin = 0; spin < 2000; spin++) {
			if (cudaMemcpy(done, e->peer_block + apo::PEER_MAX * 8, sizeof done, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); break; }
			bool all = true;
			for (int r = 0; r < e->nranks; r++) if (r != e->rank && done[r] < e->join_epoch) all = false;
			if (all) break;
			usleep(1000);
		}
	}
	for (int r = 0; r < apo::PEER_MAX; r++) {
		if (e->peer_base[r] && e->peer_ipc[r]) cudaIpcCloseMemHandle(e->peer_base[r]);
		e->peer_base[r] = nullptr; e->peer_ipc[r] = false;
	}
	if (e->peer_block) cudaFree(e->peer_block);
	e->peer_block = nullptr; e->peer_ok = false; e->peer_slot_words = 0;
}

// Allocate this rank's block, exchange handles, map the peers.  Any failure leaves peer_ok false on EVERY rank (the
// outcome is agreed with one ncclAllReduce(min)) and the join falls back to ncclAllReduce + k_finalize.
int peer_setup(apo_engine *e) {
	peer_teardown(e);
	if (e->nranks > apo::PEER_MAX || !g_nccl.AllGather) return APO_OK;
	const uint64_t slot_words = acc_words(16384, e->nranks);          // candidates up to the top-K limit fit a slot (~0.5 MB)
	const uint64_t bytes = kPeerHeader + 2 * slot_words * 8;
	PeerInfo mine{};
	mine.ok = 1;
	if (cudaMalloc((void **)&e->peer_block, bytes) != cudaSuccess) { cudaGetLastError(); e->peer_block = nullptr; mine.ok = 0; }
	if (mine.ok && cudaMemsetAsync(e->peer_block, 0, bytes, e->stream) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
	if (mine.ok && cudaIpcGetMemHandle(&mine.handle, e->peer_block) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
	mine.pid = (uint64_t)getpid(); mine.ptr = (uint64_t)(uintptr_t)e->peer_block; mine.device = e->device;
	{
		char host[256] = {0};
		gethostname(host, sizeof host - 1);
		uint64_t h = 1469598103934665603ull;
		for (const char *c = host; *c; c++) h = (h ^ (uint8_t)*c) * 1099511628211ull;
		mine.host_hash = h;
	}
	struct Scoped : DevBuf<uint8_t> { ~Scoped() { release(); } } xch;      // freed on every return path
	CK(xch.reserve(128ull * (e->nranks + 1)));
	std::vector<PeerInfo> all(e->nranks);
	CK(cudaMemcpyAsync(xch.p, &mine, 128, cudaMemcpyHostToDevice, e->stream));
	int rc = g_nccl.AllGather(xch.p, xch.p + 128, 128, kNcclInt8, e->comm, e->stream);
	if (rc != 0) { return fail(e, APO_E_NCCL, "ncclAllGather (peer handles): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); }
	CK(cudaMemcpyAsync(all.data(), xch.p + 128, 128ull * e->nranks, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	long long ok = mine.ok;
	for (int r = 0; r < e->nranks && ok; r++) {
		const PeerInfo &pi = all[r];
		if (!pi.ok || pi.host_hash != mine.host_hash) { ok = 0; break; }           // peer memory needs one NVLink domain: same box
		if (r == e->rank) { e->peer_base[r] = e->peer_block; continue; }
		if (pi.pid == mine.pid) {
			// another handle of this process (one Electron main process driving several GPUs): plain peer access
			int can = 0;
			if (cudaDeviceCanAccessPeer(&can, e->device, pi.device) != cudaSuccess || !can) { cudaGetLastError(); ok = 0; break; }
			const cudaError_t pe = cudaDeviceEnablePeerAccess(pi.device, 0);
			if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); ok = 0; break; }
			cudaGetLastError();
			e->peer_base[r] = (void *)(uintptr_t)pi.ptr;
		} else {
			void *p = nullptr;
			if (cudaIpcOpenMemHandle(&p, pi.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
			e->peer_base[r] = p; e->peer_ipc[r] = true;
		}
	}
	// touch every peer's block once now: lazily enabled peer mappings are established on first access, which must not be the first join
	if (ok) {
		for (int r = 0; r < e->nranks; r++)
			if (r != e->rank && apo::run_touch(e->peer_base[r], bytes, xch.p, e->stream) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
		if (cudaStreamSynchronize(e->stream) != cudaSuccess) { cudaGetLastError(); ok = 0; }
	}
	// agree on the outcome (also the barrier that orders every rank's zero-fill before the first join)
	long long *flag = (long long *)xch.p;
	CK(cudaMemcpyAsync(flag, &ok, 8, cudaMemcpyHostToDevice, e->stream));
	rc = g_nccl.AllReduce(flag, flag, 1, kNcclInt64, kNcclMin, e->comm, e->stream);
	if (rc != 0) { return fail(e, APO_E_NCCL, "ncclAllReduce (peer agreement): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"); }
	CK(cudaMemcpyAsync(&ok, flag, 8, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	if (!ok) { peer_teardown(e); return APO_OK; }
	e->peer_ok = true; e->peer_slot_words = slot_words; e->join_epoch = 0;
	return APO_OK;
}
}  // namespace

// =============================================================================== lifecycle
extern "C" int apo_abi_version(void) { return APO_ABI_VERSION; }

extern "C" int apo_create(int device, apo_engine **out) {
	if (!out) return fail(nullptr, APO_E_ARG, "out is NULL");
	*out = nullptr;
	int n = 0;
	cudaError_t c = cudaGetDeviceCount(&n);
	if (c != cudaSuccess || n == 0)
		return fail(nullptr, APO_E_CUDA, "no CUDA device: %s (this library has no CPU fallback)", c != cudaSuccess ? cudaGetErrorString(c) : "device count is 0");
	if (device < 0 || device >= n) return fail(nullptr, APO_E_ARG, "device %d out of range [0,%d)", device, n);
	apo_engine *e = new (std::nothrow) apo_engine();
	if (!e) return fail(nullptr, APO_E_NOMEM, "out of host memory");
	e->device = device;
	auto bail = [&](const char *what, cudaError_t ce) { fail(nullptr, APO_E_CUDA, "%s: %s", what, cudaGetErrorString(ce)); apo_destroy(e); return APO_E_CUDA; };
	if ((c = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", c);
	cudaDeviceProp prop;
	if ((c = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", c);
	if (prop.major != 10) { fail(nullptr, APO_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); apo_destroy(e); return APO_E_CUDA; }
	e->sm_count = prop.multiProcessorCount;
	e->mem_pitch_max = prop.memPitch;
	{
		// fuse heuristic (apo_score): the corpus scan rides on one warp per SM and must hide behind the evaluation stream.
		// Stream rate = 0.9 x the pin bandwidth of this device (measured on B200: 7.4 of 8.19 TB/s for Form D / R; the compact
		// layout is SM-bound at 0.61 x); scan rate = kScanRecordsPerSmCycle records per SM clock on that one warp
		// (measured: 10 M records in 3.35 ms on 148 SMs at 1.965 GHz).
		int mem_khz = 0, bus_bits = 0, sm_khz = 0;
		cudaDeviceGetAttribute(&mem_khz, cudaDevAttrMemoryClockRate, device);
		cudaDeviceGetAttribute(&bus_bits, cudaDevAttrGlobalMemoryBusWidth, device);
		cudaDeviceGetAttribute(&sm_khz, cudaDevAttrClockRate, device);
		constexpr double kStreamFracOfPin = 0.90, kCompactFracOfPin = 0.61, kScanRecordsPerSmCycle = 1.0264e-2;
		if (mem_khz > 0 && bus_bits > 0) {
			const double pin_bytes_per_ms = 2.0 * (double)mem_khz * (double)bus_bits / 8.0;     // kHz x bytes = bytes per ms
			e->stream_bytes_per_ms = kStreamFracOfPin * pin_bytes_per_ms;
			e->stream_q_bytes_per_ms = kCompactFracOfPin * pin_bytes_per_ms;
		}
		if (sm_khz > 0) e->scan_ms_per_record = 1.0 / (kScanRecordsPerSmCycle * (double)e->sm_count * (double)sm_khz);
	}
	// experiment / test switches: the environment gives the initial value once per handle (never read per call), apo_set_tuning changes them
	e->env_no_fuse = getenv("APO_NO_FUSE") != nullptr;
	e->env_force_fuse = getenv("APO_FORCE_FUSE") != nullptr;
	e->env_no_staging = getenv("APO_NO_STAGING") != nullptr;
	e->env_nccl_join = getenv("APO_JOIN_NCCL") != nullptr;
	if ((c = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", c);
	if ((c = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", c);
	e->stream = e->own_stream;
	for (auto &ev : e->ev) if ((c = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", c);
	for (int i = 0; i < 2; i++) {
		if ((c = cudaEventCreateWithFlags(&e->win_free[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", c);
		if ((c = cudaEventCreateWithFlags(&e->win_ready[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", c);
		if ((c = cudaEventCreateWithFlags(&e->stage_done[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", c);
	}
	memcpy(e->W.w, kDefaultWeights, sizeof kDefaultWeights);
	build_luts(e);
	if (upload_luts(e) != APO_OK) { g_create_error = e->err; apo_destroy(e); return APO_E_CUDA; }
	*out = e;
	return APO_OK;
}

extern "C" void apo_destroy(apo_engine *e) {
	if (!e) return;
	cudaSetDevice(e->device);
	if (e->own_stream) cudaStreamSynchronize(e->own_stream);
	peer_teardown(e);
	if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
	if (e->zc_host) cudaFreeHost(e->zc_host);
	e->q8.release(); e->qd2.release(); e->qli.release(); e->qbook.release(); e->d_ptab.release(); e->d_pair.release(); e->d_cbf.release(); e->d_d2book.release(); e->d_d2stream.release(); e->stage.release();
	e->d_lut.release(); e->corpus.release(); e->dims.release(); e->roll.release(); e->acc.release(); e->acc_joined.release(); e->acc_snapshot.release(); e->misc.release();
	e->result.release(); e->keys.release(); e->sel_key.release(); e->sel_idx.release();
	e->tups_pc.release(); e->tups_pd.release(); e->tups_d2.release();
	e->tup_l.release(); e->tup_h.release(); e->tup_pc.release(); e->tup_pd.release(); e->tup_bad.release(); e->tup_d2.release(); e->tup_ptab.release(); e->tup_val.release();
	e->win[0].release(); e->win[1].release(); e->batch_in.release(); e->batch_out.release(); e->batch_mask.release();
	if (e->h_result) cudaFreeHost(e->h_result);
	for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
	for (auto &ev : e->k1_ev) cudaEventDestroy(ev);
	for (int i = 0; i < 2; i++) { if (e->win_free[i]) cudaEventDestroy(e->win_free[i]); if (e->win_ready[i]) cudaEventDestroy(e->win_ready[i]); }
	for (int i = 0; i < 2; i++) { if (e->stage_done[i]) cudaEventDestroy(e->stage_done[i]); if (e->h_stage[i]) cudaFreeHost(e->h_stage[i]); }
	if (e->own_stream) cudaStreamDestroy(e->own_stream);
	if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
	delete e;
}

extern "C" const char *apo_last_error(const apo_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

extern "C" int apo_set_stream(apo_engine *e, uint64_t cuda_stream) {
	if (!e) return APO_E_ARG;
	e->stream = cuda_stream ? (cudaStream_t)(uintptr_t)cuda_stream : e->own_stream;
	return APO_OK;
}

extern "C" int apo_set_tuning(apo_engine *e, uint32_t flags) {
	if (!e) return APO_E_ARG;
	if (flags & ~(APO_TUNE_NO_FUSE | APO_TUNE_FORCE_FUSE | APO_TUNE_NO_STAGING | APO_TUNE_NCCL_JOIN)) return fail(e, APO_E_ARG, "unknown tuning flag %#x", flags);
	e->env_no_fuse = (flags & APO_TUNE_NO_FUSE) != 0; e->env_force_fuse = (flags & APO_TUNE_FORCE_FUSE) != 0;
	e->env_no_staging = (flags & APO_TUNE_NO_STAGING) != 0; e->env_nccl_join = (flags & APO_TUNE_NCCL_JOIN) != 0;
	return APO_OK;
}

extern "C" int apo_set_weights(apo_engine *e, const double w[APO_NDIM]) {
	if (!e || !w) return fail(e, APO_E_ARG, "NULL argument");
	for (int i = 0; i < APO_NDIM; i++)
		if (!std::isfinite(w[i]) || !(w[i] == 0.0 || (w[i] >= 1e-100 && w[i] <= 1e100)))
			return fail(e, APO_E_ARG, "weights must be 0 or within [1e-100, 1e100]");
	CK(cudaSetDevice(e->device));
	memcpy(e->W.w, w, sizeof e->W.w);
	build_luts(e);
	int rc = upload_luts(e);
	if (rc == APO_OK && e->compact) rc = upload_ptab(e);
	return rc;
}

extern "C" int apo_get_weights(const apo_engine *e, double w[APO_NDIM]) {
	if (!e || !w) return APO_E_ARG;
	memcpy(w, e->W.w, sizeof e->W.w);
	return APO_OK;
}

// =============================================================================== single-trace path
extern "C" int apo_reward_batch(apo_engine *e, const apo_record *recs, uint64_t n, double *dims, uint32_t *masks, double *finals) {
	if (!e) return APO_E_ARG;
	if (n == 0) return APO_OK;
	if (!recs) return fail(e, APO_E_ARG, "recs is NULL");
	CK(cudaSetDevice(e->device));
	constexpr uint64_t kZeroCopyMax = 1024;      // records: the IDE scores one trace at a time (TCS:408-418, 547)
	if (n <= kZeroCopyMax) {
		// single-trace path: one launch, no copies — the kernel reads the records from and writes the results to one
		// page-locked, device-mapped block (PCIe round trips of a few hundred bytes beat four cudaMemcpyAsync calls)
		const uint64_t per = sizeof(apo_record) + APO_NDIM * 8 + 8 + 4;
		if (!e->zc_host) { CK(cudaMallocHost((void **)&e->zc_host, kZeroCopyMax * per)); e->zc_cap = kZeroCopyMax * per; }
		apo_record *zin = (apo_record *)e->zc_host;
		double *zdims = (double *)(e->zc_host + kZeroCopyMax * sizeof(apo_record));
		double *zfin = zdims + kZeroCopyMax * APO_NDIM;
		uint32_t *zmask = (uint32_t *)(zfin + kZeroCopyMax);
		memcpy(zin, recs, n * sizeof(apo_record));
		CK(apo::run_reward_batch(zin, n, e->W, e->d_lut.p, zdims, zmask, zfin, e->stream));
		CK(cudaStreamSynchronize(e->stream));
		if (dims) memcpy(dims, zdims, n * APO_NDIM * 8);
		if (masks) memcpy(masks, zmask, n * 4);
		if (finals) memcpy(finals, zfin, n * 8);
		return APO_OK;
	}
	CK(e->batch_in.reserve(n));
	CK(e->batch_out.reserve(n * (APO_NDIM + 1)));
	CK(e->batch_mask.reserve(n));
	CK(cudaMemcpyAsync(e->batch_in.p, recs, n * sizeof(apo_record), cudaMemcpyHostToDevice, e->stream));
	CK(apo::run_reward_batch(e->batch_in.p, n, e->W, e->d_lut.p, e->batch_out.p, e->batch_mask.p, e->batch_out.p + n * APO_NDIM, e->stream));
	if (dims) CK(cudaMemcpyAsync(dims, e->batch_out.p, n * APO_NDIM * 8, cudaMemcpyDeviceToHost, e->stream));
	if (masks) CK(cudaMemcpyAsync(masks, e->batch_mask.p, n * 4, cudaMemcpyDeviceToHost, e->stream));
	if (finals) CK(cudaMemcpyAsync(finals, e->batch_out.p + n * APO_NDIM, n * 8, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return APO_OK;
}

extern "C" int apo_reward_one(apo_engine *e, const apo_record *rec, double dims[APO_NDIM], uint32_t *mask, double *final_reward) {
	return apo_reward_batch(e, rec, 1, dims, mask, final_reward);
}

// =============================================================================== corpus
extern "C" int apo_corpus_upload(apo_engine *e, const apo_record *recs, uint64_t T, uint64_t idx_base) {
	if (!e) return APO_E_ARG;
	if (T && !recs) return fail(e, APO_E_ARG, "recs is NULL");
	CK(cudaSetDevice(e->device));
	CK(e->corpus.reserve(T ? T : 1));
	if (T) { const int rc = h2d_rows(e, e->corpus.p, T * sizeof(apo_record), recs, T * sizeof(apo_record), T * sizeof(apo_record), 1, e->stream); if (rc) return rc; }
	CK(apo::run_fill_durclass(e->corpus.p, 32, T, e->stream));        // resident records always carry the duration class
	CK(cudaStreamSynchronize(e->stream));
	e->corpus_T = T; e->corpus_base = idx_base;
	return APO_OK;
}

extern "C" int apo_corpus_upload_json(apo_engine *e, const char *json, uint64_t len, uint64_t idx_base, uint64_t *n_records) {
	if (!e) return APO_E_ARG;
	if (!json) return fail(e, APO_E_ARG, "json is NULL");
	std::vector<apo_record> recs;
	uint64_t pos = 0;
	int64_t n;
	try {
		recs.resize(len / 256 + 16);                      // a persisted trace is never shorter than ~300 bytes
		n = apo_records_from_json(json, len, recs.data(), recs.size(), &pos);
		if (n >= 0 && uint64_t(n) > recs.size()) {
			recs.resize(size_t(n));
			n = apo_records_from_json(json, len, recs.data(), recs.size(), &pos);
		}
	} catch (const std::bad_alloc &) {
		return fail(e, APO_E_NOMEM, "host allocation for %llu bytes of JSON failed", (unsigned long long)len);
	}
	if (n < 0) return fail(e, APO_E_ARG, "malformed trace JSON at byte %llu", (unsigned long long)pos);
	if (n_records) *n_records = uint64_t(n);
	return apo_corpus_upload(e, recs.data(), uint64_t(n), idx_base);
}

extern "C" int apo_corpus_generate(apo_engine *e, uint64_t seed, uint64_t t0, uint64_t T, uint32_t agent_permille) {
	if (!e) return APO_E_ARG;
	CK(cudaSetDevice(e->device));
	CK(e->corpus.reserve(T ? T : 1));
	CK(apo::run_gen_records(e->corpus.p, T, seed, 1u, 0, 1, t0, T, agent_permille, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	e->corpus_T = T; e->corpus_base = t0;
	return APO_OK;
}

extern "C" int apo_corpus_download(apo_engine *e, apo_record *out, uint64_t first, uint64_t n) {
	if (!e || !out) return fail(e, APO_E_ARG, "NULL argument");
	if (first + n > e->corpus_T) return fail(e, APO_E_ARG, "range outside the corpus");
	CK(cudaSetDevice(e->device));
	CK(cudaMemcpyAsync(out, e->corpus.p + first, n * sizeof(apo_record), cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return APO_OK;
}

// =============================================================================== evaluations
extern "C" int apo_dims_upload(apo_engine *e, const float *dims, uint32_t C, uint64_t T) {
	if (!e) return APO_E_ARG;
	if (C && T && !dims) return fail(e, APO_E_ARG, "dims is NULL");
	CK(cudaSetDevice(e->device));
	const uint64_t pitch = round_up(T ? T : 1, 32);
	CK(e->dims.reserve((uint64_t)(C ? C : 1) * pitch * APO_NDIM));
	if (C && T) {
		// pad evaluations of every row are all-NaN = finalReward null
		if (pitch != T) CK(cudaMemsetAsync(e->dims.p, 0xFF, (uint64_t)C * pitch * APO_NDIM * 4, e->stream));
		{ const int rc = h2d_rows(e, e->dims.p, pitch * APO_NDIM * 4, dims, T * APO_NDIM * 4, T * APO_NDIM * 4, C, e->stream); if (rc) return rc; }
	}
	CK(cudaStreamSynchronize(e->stream));
	e->dims_ptr = e->dims.p; e->dims_C = C; e->dims_T = T; e->dims_pitch = pitch; e->compact = false;
	return APO_OK;
}

extern "C" int apo_dims_generate(apo_engine *e, uint64_t seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T, uint32_t agent_permille) {
	if (!e) return APO_E_ARG;
	CK(cudaSetDevice(e->device));
	const uint64_t pitch = round_up(T ? T : 1, 32);
	CK(e->dims.reserve((uint64_t)(C ? C : 1) * pitch * APO_NDIM));
	if (pitch != T && C) CK(cudaMemsetAsync(e->dims.p, 0xFF, (uint64_t)C * pitch * APO_NDIM * 4, e->stream));
	CK(apo::run_gen_dims(e->dims.p, pitch, seed, c0, C, t0, T, agent_permille, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	e->dims_ptr = e->dims.p; e->dims_C = C; e->dims_T = T; e->dims_pitch = pitch; e->compact = false;
	return APO_OK;
}

extern "C" int apo_dims_download(apo_engine *e, float *out, uint32_t c, uint64_t first, uint64_t n) {
	if (!e || !out) return fail(e, APO_E_ARG, "NULL argument");
	if ((!e->dims_ptr && !e->compact) || c >= e->dims_C || first + n > e->dims_T) return fail(e, APO_E_ARG, "range outside the evaluations");
	CK(cudaSetDevice(e->device));
	if (e->compact) {
		CK(e->stage.reserve((n ? n : 1) * APO_NDIM));
		const uint64_t off = (uint64_t)c * e->dims_pitch + first;
		CK(apo::run_decode(e->q8.p + off, e->qd2.p + off, e->qli.p + off, n, e->qbook.p, e->stage.p, e->stream));
		CK(cudaMemcpyAsync(out, e->stage.p, n * APO_NDIM * 4, cudaMemcpyDeviceToHost, e->stream));
		CK(cudaStreamSynchronize(e->stream));
		return APO_OK;
	}
	CK(cudaMemcpyAsync(out, e->dims_ptr + ((uint64_t)c * e->dims_pitch + first) * APO_NDIM, n * APO_NDIM * 4, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return APO_OK;
}

extern "C" int apo_dims_attach(apo_engine *e, uint64_t device_ptr, uint32_t C, uint64_t T, uint64_t pitch_evals) {
	if (!e) return APO_E_ARG;
	if (!device_ptr || device_ptr % 16) return fail(e, APO_E_ARG, "device pointer must be non-NULL and 16-byte aligned");
	if (pitch_evals % 4 || pitch_evals < round_up(T, 4)) return fail(e, APO_E_ARG, "pitch must be a multiple of 4 evaluations and >= T rounded up to 4");
	e->dims_ptr = (const float *)(uintptr_t)device_ptr; e->dims_C = C; e->dims_T = T; e->dims_pitch = pitch_evals; e->compact = false;
	return APO_OK;
}

extern "C" int apo_dims_layout(const apo_engine *e) {
	if (!e) return 0;
	if (e->compact) return 2;
	return (e->dims_ptr && e->dims_C) ? 1 : 0;
}

extern "C" int apo_dims_compact(apo_engine *e) {
	if (!e) return APO_E_ARG;
	if (e->compact) return APO_OK;
	if (!e->dims_ptr || e->dims_C == 0) return fail(e, APO_E_STATE, "no dims loaded");
	CK(cudaSetDevice(e->device));
	const uint32_t C = e->dims_C; const uint64_t T = e->dims_T, pitch_in = e->dims_pitch, pitch = round_up(T ? T : 1, 32);
	int rc = compact_begin(e, C, pitch);
	if (rc) return rc;
	CK(apo::run_transcode(e->dims_ptr, pitch_in, C, T, e->q8.p, e->qd2.p, pitch, e->qbook.p, e->qbook.p + 8 * 256, e->stream));
	const float *keep = e->dims_ptr;
	if ((rc = compact_finish(e, C, T, pitch))) { e->dims_ptr = keep; return rc; }
	e->dims.release();                                  // the fp32 copy (if engine-owned) is no longer needed
	return APO_OK;
}

namespace {
// chunked producer -> transcoder: at most ~2 GB of fp32 staging, never the whole tensor
template <class Fill>
int compact_stream(apo_engine *e, uint32_t C, uint64_t T, Fill fill) {
	CK(cudaSetDevice(e->device));
	const uint64_t pitch = round_up(T ? T : 1, 32);
	uint64_t Cc = (2ull << 30) / (pitch * 36);
	if (Cc < 1) Cc = 1;
	if (Cc > C) Cc = C ? C : 1;
	CK(e->stage.reserve(Cc * pitch * APO_NDIM));
	int rc = compact_begin(e, C, pitch);
	if (rc) return rc;
	for (uint32_t c0 = 0; c0 < C; c0 += (uint32_t)Cc) {
		const uint32_t cn = C - c0 < Cc ? C - c0 : (uint32_t)Cc;
		if ((rc = fill(c0, cn, e->stage.p, pitch))) return rc;
		CK(apo::run_transcode(e->stage.p, pitch, cn, T, e->q8.p + (uint64_t)c0 * pitch, e->qd2.p + (uint64_t)c0 * pitch, pitch,
		                      e->qbook.p, e->qbook.p + 8 * 256, e->stream));
	}
	e->compact = false; e->dims_ptr = nullptr; e->dims_C = 0;          // whatever was loaded before is superseded
	rc = compact_finish(e, C, T, pitch);
	e->stage.release();
	if (rc == APO_OK) e->dims.release();
	return rc;
}
}  // namespace

extern "C" int apo_dims_generate_compact(apo_engine *e, uint64_t seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T, uint32_t agent_permille) {
	if (!e) return APO_E_ARG;
	return compact_stream(e, C, T, [&](uint32_t cb, uint32_t cn, float *stage, uint64_t pitch) -> int {
		CK(apo::run_gen_dims(stage, pitch, seed, c0 + cb, cn, t0, T, agent_permille, e->stream));
		return APO_OK;
	});
}

extern "C" int apo_dims_upload_compact(apo_engine *e, const float *dims, uint32_t C, uint64_t T) {
	if (!e) return APO_E_ARG;
	if (C && T && !dims) return fail(e, APO_E_ARG, "dims is NULL");
	return compact_stream(e, C, T, [&](uint32_t cb, uint32_t cn, float *stage, uint64_t pitch) -> int {
		const int rc = h2d_rows(e, stage, pitch * APO_NDIM * 4, dims + (uint64_t)cb * T * APO_NDIM, T * APO_NDIM * 4, T * APO_NDIM * 4, cn, e->stream);
		if (rc) return rc;
		CK(cudaStreamSynchronize(e->stream));              // the host chunk may be reused by the caller after return
		return APO_OK;
	});
}

namespace {
int rollouts_upload_rows(apo_engine *e, const void *recs, uint32_t row, uint32_t C, uint64_t T) {
	if (!e) return APO_E_ARG;
	if (C && T && !recs) return fail(e, APO_E_ARG, "recs is NULL");
	CK(cudaSetDevice(e->device));
	const uint64_t pitch = round_up(T ? T : 1, 4);
	CK(e->roll.reserve((uint64_t)(C ? C : 1) * pitch * row));
	if (C && T) {
		if (pitch != T) CK(cudaMemsetAsync(e->roll.p, 0, (uint64_t)C * pitch * row, e->stream));   // VALID clear
		{ const int rc = h2d_rows(e, e->roll.p, pitch * row, recs, T * row, T * row, C, e->stream); if (rc) return rc; }
		CK(apo::run_fill_durclass(e->roll.p, row, (uint64_t)C * pitch, e->stream));
	}
	CK(cudaStreamSynchronize(e->stream));
	e->roll_C = C; e->roll_T = T; e->roll_pitch = pitch; e->roll_row = row;
	return APO_OK;
}
int rollouts_generate_rows(apo_engine *e, uint32_t row, uint64_t seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T, uint32_t agent_permille) {
	if (!e) return APO_E_ARG;
	CK(cudaSetDevice(e->device));
	const uint64_t pitch = round_up(T ? T : 1, 4);
	CK(e->roll.reserve((uint64_t)(C ? C : 1) * pitch * row));
	if (pitch != T && C) CK(cudaMemsetAsync(e->roll.p, 0, (uint64_t)C * pitch * row, e->stream));
	if (row == 32) CK(apo::run_gen_records((apo_record *)e->roll.p, pitch, seed, 2u, c0, C, t0, T, agent_permille, e->stream));
	else CK(apo::run_gen_records16((apo_record16 *)e->roll.p, pitch, seed, 2u, c0, C, t0, T, agent_permille, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	e->roll_C = C; e->roll_T = T; e->roll_pitch = pitch; e->roll_row = row;
	return APO_OK;
}
int rollouts_download_rows(apo_engine *e, void *out, uint32_t row, uint32_t c, uint64_t first, uint64_t n) {
	if (!e || !out) return fail(e, APO_E_ARG, "NULL argument");
	if (!e->roll.p || e->roll_row != row) return fail(e, APO_E_STATE, "no rollouts of that row size loaded");
	if (c >= e->roll_C || first + n > e->roll_T) return fail(e, APO_E_ARG, "range outside the rollouts");
	CK(cudaSetDevice(e->device));
	CK(cudaMemcpyAsync(out, e->roll.p + ((uint64_t)c * e->roll_pitch + first) * row, n * row, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return APO_OK;
}
}  // namespace

extern "C" int apo_rollouts_upload(apo_engine *e, const apo_record *recs, uint32_t C, uint64_t T) { return rollouts_upload_rows(e, recs, 32, C, T); }
extern "C" int apo_rollouts16_upload(apo_engine *e, const apo_record16 *recs, uint32_t C, uint64_t T) { return rollouts_upload_rows(e, recs, 16, C, T); }
extern "C" int apo_rollouts_generate(apo_engine *e, uint64_t seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T, uint32_t agent_permille) {
	return rollouts_generate_rows(e, 32, seed, c0, C, t0, T, agent_permille);
}
extern "C" int apo_rollouts16_generate(apo_engine *e, uint64_t seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T, uint32_t agent_permille) {
	return rollouts_generate_rows(e, 16, seed, c0, C, t0, T, agent_permille);
}
extern "C" int apo_rollouts_download(apo_engine *e, apo_record *out, uint32_t c, uint64_t first, uint64_t n) { return rollouts_download_rows(e, out, 32, c, first, n); }
extern "C" int apo_rollouts16_download(apo_engine *e, apo_record16 *out, uint32_t c, uint64_t first, uint64_t n) { return rollouts_download_rows(e, out, 16, c, first, n); }

extern "C" int apo_record_pack16(const apo_record *in, uint64_t n, apo_record16 *out, uint64_t *bad) {
	if ((!in || !out) && n) return APO_E_ARG;
	for (uint64_t i = 0; i < n; i++) {
		if (!apo::representable16(in[i])) { if (bad) *bad = i; return APO_E_ARG; }
		out[i] = apo::pack16(in[i]);
	}
	return APO_OK;
}
extern "C" int apo_record_unpack16(const apo_record16 *in, uint64_t n, apo_record *out) {
	if ((!in || !out) && n) return APO_E_ARG;
	for (uint64_t i = 0; i < n; i++) out[i] = apo::unpack16(in[i]);
	return APO_OK;
}

// =============================================================================== scoring
static int source_shape(apo_engine *e, const apo_score_opts *o, uint32_t *C, uint64_t *T) {
	if (!o) return fail(e, APO_E_ARG, "opts is NULL");
	if (o->source > APO_SRC_TUPLES) return fail(e, APO_E_ARG, "unknown source %u", o->source);
	if (o->source == APO_SRC_TUPLES) {
		if (e->tup_l.p == nullptr || e->tup_C == 0) return fail(e, APO_E_STATE, "no tuples loaded");
		*C = e->tup_C; *T = e->tup_T;
		return APO_OK;
	}
	const bool raw = o->source == APO_SRC_ROLLOUTS;
	if (raw ? (e->roll.p == nullptr || e->roll_C == 0) : ((e->dims_ptr == nullptr && !e->compact) || e->dims_C == 0))
		return fail(e, APO_E_STATE, "no %s loaded", raw ? "rollouts" : "dims");
	*C = raw ? e->roll_C : e->dims_C;
	*T = raw ? e->roll_T : e->dims_T;
	return APO_OK;
}

extern "C" int apo_score_begin(apo_engine *e, uint32_t C_total) {
	if (!e) return APO_E_ARG;
	if (C_total == 0) return fail(e, APO_E_ARG, "C_total is 0");
	CK(cudaSetDevice(e->device));
	int rc = ensure_scratch(e, C_total, C_total);
	if (rc) return rc;
	e->timing_on = true;
	return begin_score(e, C_total);
}

extern "C" int apo_score_accumulate(apo_engine *e, const apo_score_opts *o, uint32_t cand_offset) {
	if (!e) return APO_E_ARG;
	if (!e->scoring) return fail(e, APO_E_STATE, "apo_score_begin has not been called");
	uint32_t C = 0; uint64_t T = 0, first = 0, count = 0;
	int rc = source_shape(e, o, &C, &T);
	if (rc) return rc;
	if ((uint64_t)cand_offset + C > e->score_C) return fail(e, APO_E_ARG, "candidates [%u,%u) exceed C_total=%u", cand_offset, cand_offset + C, e->score_C);
	if (o->first % 4) return fail(e, APO_E_ARG, "window start must be a multiple of 4");
	if (o->first > T || (o->count && o->first + o->count > T)) return fail(e, APO_E_ARG, "window outside [0,%llu)", (unsigned long long)T);
	first = o->first; count = o->count ? o->count : T - o->first;
	CK(cudaSetDevice(e->device));
	return launch_k1_resident(e, o, cand_offset, first, count);
}

extern "C" int apo_score_finish(apo_engine *e, const apo_score_opts *o, double *scores, uint64_t *counts, int32_t *topk,
                                apo_corpus_report *report) {
	if (!e) return APO_E_ARG;
	if (!o) return fail(e, APO_E_ARG, "opts is NULL");
	if (!e->scoring) return fail(e, APO_E_STATE, "apo_score_begin has not been called");
	if (o->K > e->score_C) return fail(e, APO_E_ARG, "K=%u exceeds the number of candidates C=%u", o->K, e->score_C);
	if (o->K > kMaxK) return fail(e, APO_E_ARG, "K=%u exceeds the supported beam width %u", o->K, kMaxK);
	CK(cudaSetDevice(e->device));
	if (peer_join_active(e, e->score_C)) e->join_epoch++;
	return finish_score(e, o, e->score_C, scores, counts, topk, report);
}

extern "C" int apo_score(apo_engine *e, const apo_score_opts *o, double *scores, uint64_t *counts, int32_t *topk, apo_corpus_report *report) {
	if (!e) return APO_E_ARG;
	uint32_t C = 0; uint64_t T = 0, first = 0, count = 0;
	int rc = source_shape(e, o, &C, &T);
	if (rc) return rc;
	if ((rc = check_opts(e, o, C, T, &first, &count))) return rc;
	CK(cudaSetDevice(e->device));
	if ((rc = ensure_scratch(e, C, o->K))) return rc;
	const bool tup = o->source == APO_SRC_TUPLES;      // gather-and-add kernel: stand-alone corpus scan + tail
	const double row_bytes = tup ? 3.0 : o->source == APO_SRC_ROLLOUTS ? (double)e->roll_row : (e->compact ? 14.0 : 36.0);
	const double stream_bytes = (double)C * (double)count * row_bytes;
	choose_timing(e, o, (uint64_t)stream_bytes);
	if (peer_join_active(e, C)) e->join_epoch++;          // counts joined calls only: consecutive joins alternate between the two export slots
	// One launch per call: the corpus scan (K2) rides on an extra warp of every scoring CTA and the last CTA finalises (K3,
	// after the peer-memory join when the record axis is sharded) ... when the scan can hide behind the evaluation stream.
	// Both rates are derived from the device at apo_create.
	const double k1_ms = stream_bytes / (o->source == APO_SRC_DIMS && e->compact ? e->stream_q_bytes_per_ms : e->stream_bytes_per_ms);
	const double scan_ms = (double)e->corpus_T * e->scan_ms_per_record;
	const bool one_launch = e->nranks == 1 || peer_join_active(e, C);
	// ... or when it is tiny anyway (the IDE's real corpora, <= 1000 traces): at most 16 records per lane of the corpus warps,
	// a few microseconds, against a second launch
	const int tile_evals = tup ? apo::kt_tile_evals() : (o->source == APO_SRC_DIMS && e->compact) ? apo::kq_tile_evals((int)o->variant) : apo::k1_tile_evals((int)row_bytes, (int)o->variant);
	const uint64_t tiles = (uint64_t)C * ((count + (uint64_t)tile_evals - 1) / (uint64_t)tile_evals);
	const uint64_t ctas = tiles < (uint64_t)e->sm_count ? tiles : (uint64_t)e->sm_count;
	const bool small_scan = e->corpus_T <= 16ull * 32ull * ctas;
	const bool fuse = !tup && wants_corpus(e, o) && count > 0 && !e->env_no_fuse && (k1_ms > 1.3 * scan_ms || small_scan || e->env_force_fuse);
	// without a corpus request the same tail still saves the K3 launch: an empty scan, then the last CTA finalises
	const bool tail_only = !tup && !wants_corpus(e, o) && one_launch && count > 0 && !e->env_no_fuse;
	if (fuse || tail_only) {
		if ((rc = begin_score(e, C, true))) return rc;
		apo::FinalizeParams F = make_fin(e, C, o->K, fuse ? 1 : 0);
		const bool self_clean = one_launch;          // the finalising CTA is in this launch: it can leave the block zeroed
		if (self_clean) { F.clean_ptr = e->acc.p; F.clean_words = (uint32_t)(acc_words(C, e->nranks) + ARM_WORDS); F.snapshot = e->acc_snapshot.p; }
		apo::K2Params k2 = make_k2(e, C, F);
		if (tail_only) { k2.T = 0; k2.recs = nullptr; }
		if ((rc = launch_k1_resident(e, o, 0, first, count, &k2))) return rc;
		rc = finish_score(e, o, C, scores, counts, topk, report, true);
		if (rc == APO_OK && self_clean) {
			e->acc_clean_words = F.clean_words;
			if (e->nranks == 1) e->partials_src = e->acc_snapshot.p;
		}
		return rc;
	} else {
		if ((rc = begin_score(e, C))) return rc;
		if ((rc = launch_k1_resident(e, o, 0, first, count))) return rc;
	}
	return finish_score(e, o, C, scores, counts, topk, report, fuse || tail_only);
}

namespace {
// Streams host rows [C][T] (row bytes 36 = Form D, 32 = Form R, 16 = Form R16) through two device
// windows: the H2D copy of chunk i+1 (copy stream) overlaps K1 on chunk i (compute stream).
int score_host_rows_impl(apo_engine *e, const apo_score_opts *o, const uint8_t *rows, uint32_t row, uint32_t C, uint64_t T,
                         double *scores, uint64_t *counts, int32_t *topk, apo_corpus_report *report) {
	if (!e) return APO_E_ARG;
	if (!o) return fail(e, APO_E_ARG, "opts is NULL");
	if (!rows || C == 0) return fail(e, APO_E_ARG, "input is NULL or C == 0");
	if (o->K > C) return fail(e, APO_E_ARG, "K=%u exceeds the number of candidates C=%u", o->K, C);
	if (o->K > kMaxK) return fail(e, APO_E_ARG, "K=%u exceeds the supported beam width %u", o->K, kMaxK);
	if (o->first || o->count) return fail(e, APO_E_ARG, "windows are not supported by the host-streaming calls");
	CK(cudaSetDevice(e->device));
	int rc;
	if ((rc = ensure_scratch(e, C, o->K))) return rc;
	// window of ~256 MB per buffer, a multiple of the K1 tile so chunks never split a tile
	const int tile = apo::k1_tile_evals((int)row, (int)o->variant);
	uint64_t Tc = (256ull << 20) / ((uint64_t)C * row);
	Tc = Tc / tile * tile;
	if (Tc < (uint64_t)tile) Tc = tile;
	if (Tc > round_up(T ? T : 1, tile)) Tc = round_up(T ? T : 1, tile);
	for (int i = 0; i < 2; i++) CK(e->win[i].reserve((uint64_t)C * Tc * row));
	// pageable callers (malloc / ArrayBuffer / numpy): the driver would stage such copies through one small internal
	// buffer on the calling thread; instead gather each chunk into pinned memory with a few host threads while the
	// previous chunk is on the wire.  Pinned or registered callers (apo_host_alloc) are read in place.
	const bool staged = T > 0 && is_pageable(rows) && !e->env_no_staging;
	if (staged && (rc = ensure_stage(e, (uint64_t)C * Tc * row))) return rc;
	choose_timing(e, o, (uint64_t)C * T * row);
	if (peer_join_active(e, C)) e->join_epoch++;
	if ((rc = begin_score(e, C))) return rc;
	const bool recip = (o->flags & APO_SCORE_RECIP) != 0;
	int nchunk = 0;
	for (uint64_t t0 = 0; t0 < T; t0 += Tc, nchunk++) {
		const int b = nchunk & 1;
		const uint64_t n = T - t0 < Tc ? T - t0 : Tc;
		const uint8_t *src = rows + t0 * row;
		size_t spitch = T * row;
		if (staged) {
			if (nchunk >= 2) CK(cudaEventSynchronize(e->stage_done[b]));      // the copy that read this staging buffer has finished
			parallel_rows_copy(e->h_stage[b], Tc * row, src, spitch, n * row, C);
			src = e->h_stage[b]; spitch = Tc * row;
		}
		if (nchunk >= 2) CK(cudaStreamWaitEvent(e->copy_stream, e->win_free[b], 0));
		if ((rc = copy_rows_h2d(e, e->win[b].p, Tc * row, src, spitch, n * row, C, e->copy_stream))) { cudaStreamSynchronize(e->copy_stream); return rc; }
		if (staged) CK(cudaEventRecord(e->stage_done[b], e->copy_stream));
		CK(cudaEventRecord(e->win_ready[b], e->copy_stream));
		CK(cudaStreamWaitEvent(e->stream, e->win_ready[b], 0));
		if (row != 36) {                                               // trace records: make sure the window carries the duration class
			CK(apo::run_fill_durclass(e->win[b].p, row, (uint64_t)C * Tc, e->stream));
			e->timing.launches++;
		}
		apo::K1Params P{};
		P.base = e->win[b].p; P.pitch_bytes = Tc * row; P.C = C; P.T = n; P.acc = e->acc.p;
		P.lut = e->d_lut.p; P.W = e->W;
		if ((rc = record_k1_event(e, 0))) return rc;
		CK(apo::run_reward9(P, (int)row, (int)o->variant, recip, e->sm_count, e->stream));
		if ((rc = record_k1_event(e, 1))) return rc;
		e->timing.launches++;
		CK(cudaEventRecord(e->win_free[b], e->stream));
	}
	return finish_score(e, o, C, scores, counts, topk, report);
}

int score_host_rows(apo_engine *e, const apo_score_opts *o, const uint8_t *rows, uint32_t row, uint32_t C, uint64_t T,
                    double *scores, uint64_t *counts, int32_t *topk, apo_corpus_report *report) {
	const int rc = score_host_rows_impl(e, o, rows, row, C, T, scores, counts, topk, report);
	if (rc && e) {
		// a failure in the middle of the chunk loop: nothing may still be reading the caller's memory when we return
		cudaStreamSynchronize(e->copy_stream);
		cudaStreamSynchronize(e->stream);
	}
	return rc;
}
}  // namespace

// =============================================================================== Form Q in host memory (compact wire format)
extern "C" int apo_dims_codebook(apo_engine *e, uint32_t *codebook) {
	if (!e || !codebook) return fail(e, APO_E_ARG, "NULL argument");
	if (!e->compact) return fail(e, APO_E_STATE, "no compact (Form Q) evaluations loaded");
	memcpy(codebook, e->qbook_host, sizeof e->qbook_host);
	return APO_OK;
}

extern "C" int apo_dims_compact_download(apo_engine *e, uint64_t *q8, float *d2, uint16_t *li, uint32_t c, uint64_t first, uint64_t n) {
	if (!e || !q8 || !d2 || !li) return fail(e, APO_E_ARG, "NULL argument");
	if (!e->compact) return fail(e, APO_E_STATE, "no compact (Form Q) evaluations loaded");
	if (c >= e->dims_C || first + n > e->dims_T) return fail(e, APO_E_ARG, "range outside the evaluations");
	CK(cudaSetDevice(e->device));
	const uint64_t off = (uint64_t)c * e->dims_pitch + first;
	CK(cudaMemcpyAsync(q8, e->q8.p + off, n * 8, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaMemcpyAsync(d2, e->qd2.p + off, n * 4, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaMemcpyAsync(li, e->qli.p + off, n * 2, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return APO_OK;
}

namespace {
// d2book of the resident Form Q tensor: distinct tool_success_rate values ordered by value (built once per loaded tensor)
int ensure_d2book(apo_engine *e) {
	if (e->d2book_n != ~0u) return APO_OK;
	DevBuf<uint32_t> table;
	CK(table.reserve(16384 + 1));
	CK(cudaMemsetAsync(table.p, 0xFF, 16384 * 4, e->stream));
	CK(cudaMemsetAsync(table.p + 16384, 0, 4, e->stream));
	// pad evaluations carry a cleared presence bit, so the whole pitched planes can be scanned
	CK(apo::run_collect_d2(e->qd2.p, e->qli.p, (uint64_t)e->dims_C * e->dims_pitch, table.p, table.p + 16384, e->stream));
	std::vector<uint32_t> host(16384 + 1);
	CK(cudaMemcpyAsync(host.data(), table.p, host.size() * 4, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	table.release();
	if (host[16384] > 4095) return fail(e, APO_E_STATE, "tool_success_rate takes %u distinct values: more than the 4095 Form P can index", host[16384]);
	std::vector<uint32_t> vals;
	for (int i = 0; i < 16384; i++) if (host[i] != 0xFFFFFFFFu) vals.push_back(host[i]);
	std::sort(vals.begin(), vals.end(), [](uint32_t x, uint32_t y) {
		float fx, fy; memcpy(&fx, &x, 4); memcpy(&fy, &y, 4);
		if (fx != fy) return fx < fy;
		return x > y;
	});
	for (size_t c = 0; c < 4096; c++) e->d2book_host[c] = c < vals.size() ? vals[c] : 0xFFFFFFFFu;
	CK(e->d_d2book.reserve(4096));
	CK(cudaMemcpyAsync(e->d_d2book.p, e->d2book_host, 4096 * 4, cudaMemcpyHostToDevice, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	e->d2book_n = (uint32_t)vals.size();
	return APO_OK;
}
}  // namespace

extern "C" int apo_dims_d2book(apo_engine *e, uint32_t *d2book) {
	if (!e || !d2book) return fail(e, APO_E_ARG, "NULL argument");
	if (!e->compact) return fail(e, APO_E_STATE, "no compact (Form Q) evaluations loaded");
	CK(cudaSetDevice(e->device));
	const int rc = ensure_d2book(e);
	if (rc) return rc;
	memcpy(d2book, e->d2book_host, sizeof e->d2book_host);
	return APO_OK;
}

extern "C" int apo_dims_packed_download(apo_engine *e, uint32_t *pc, uint16_t *pd, uint32_t c, uint64_t first, uint64_t n) {
	if (!e || !pc || !pd) return fail(e, APO_E_ARG, "NULL argument");
	if (!e->compact) return fail(e, APO_E_STATE, "no compact (Form Q) evaluations loaded");
	if (c >= e->dims_C || first + n > e->dims_T) return fail(e, APO_E_ARG, "range outside the evaluations");
	CK(cudaSetDevice(e->device));
	int rc = ensure_d2book(e);
	if (rc) return rc;
	if (n == 0) return APO_OK;
	CK(e->stage.reserve((n * 6 + 3) / 4 + 16));
	uint32_t *dpc = (uint32_t *)e->stage.p;
	unsigned short *dpd = (unsigned short *)(dpc + n);
	uint32_t *bad = (uint32_t *)(e->stage.p + (n * 6 + 3) / 4 + 8);
	CK(cudaMemsetAsync(bad, 0, 4, e->stream));
	const uint64_t off = (uint64_t)c * e->dims_pitch + first;
	CK(apo::run_pack_p(e->q8.p + off, e->qd2.p + off, e->qli.p + off, n, (const uint32_t *)e->d_d2book.p, e->d2book_n, dpc, dpd, bad, e->stream));
	uint32_t hbad = 0;
	CK(cudaMemcpyAsync(pc, dpc, n * 4, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaMemcpyAsync(pd, dpd, n * 2, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	if (hbad) return fail(e, APO_E_STATE, "a coded dimension takes more than 15 distinct values: not representable in Form P");
	return APO_OK;
}

namespace {
// Streams the three Form Q planes [C][T] (8 + 4 + 2 bytes per evaluation) through two device windows, H2D of chunk i+1
// overlapped with K1q on chunk i — the host-streaming call for callers that hold the compact wire format (14 B / evaluation
// over PCIe instead of 36).
// Form P (pc / pd / d2book non-NULL): 6 B per evaluation cross PCIe and k_unpack_p expands each window to Form Q on the device.
int score_host_compact_impl(apo_engine *e, const apo_score_opts *o, const uint64_t *q8, const float *d2, const uint16_t *li,
                            const uint32_t *pc, const uint16_t *pd, const uint32_t *d2book,
                            const uint32_t *codebook, uint32_t C, uint64_t T, double *scores, uint64_t *counts, int32_t *topk,
                            apo_corpus_report *report) {
	const bool packed = pc != nullptr;
	if (!o) return fail(e, APO_E_ARG, "opts is NULL");
	if (!codebook || C == 0 || (packed ? (!pd || !d2book) : (!q8 || !d2 || !li))) return fail(e, APO_E_ARG, "input is NULL or C == 0");
	if (o->K > C) return fail(e, APO_E_ARG, "K=%u exceeds the number of candidates C=%u", o->K, C);
	if (o->K > kMaxK) return fail(e, APO_E_ARG, "K=%u exceeds the supported beam width %u", o->K, kMaxK);
	if (o->first || o->count) return fail(e, APO_E_ARG, "windows are not supported by the host-streaming calls");
	CK(cudaSetDevice(e->device));
	int rc;
	if ((rc = ensure_scratch(e, C, o->K))) return rc;
	// the lookup tables of K1q follow the codebook of THIS call; a resident Form Q tensor keeps its own and gets it back when the
	// call returns, on every path (Restore)
	struct Restore {
		apo_engine *e; uint32_t saved[8 * 256]; bool armed = false;
		~Restore() { if (armed) { memcpy(e->qbook_host, saved, sizeof saved); upload_ptab(e); } }
	} restore{e};
	const bool same_book = memcmp(e->qbook_host, codebook, sizeof e->qbook_host) == 0 && e->d_ptab.p != nullptr;
	if (!same_book) {
		if (e->compact) { memcpy(restore.saved, e->qbook_host, sizeof restore.saved); restore.armed = true; }
		memcpy(e->qbook_host, codebook, sizeof e->qbook_host);
		if ((rc = upload_ptab(e))) return rc;
	}
	const int qv = e->q_pair_ok ? 0 : 4;
	const int tile = apo::kq_tile_evals(qv);
	const uint64_t wire = packed ? 6 : 14;
	uint64_t Tc = (256ull << 20) / ((uint64_t)C * wire);
	Tc = Tc / tile * tile;
	if (Tc < (uint64_t)tile) Tc = tile;
	if (Tc > round_up(T ? T : 1, tile)) Tc = round_up(T ? T : 1, tile);
	for (int i = 0; i < 2; i++) CK(e->win[i].reserve((uint64_t)C * Tc * (packed ? 20 : 14)));
	if (packed) {
		CK(e->d_d2stream.reserve(4096));
		CK(cudaMemcpyAsync(e->d_d2stream.p, d2book, 4096 * 4, cudaMemcpyHostToDevice, e->stream));   // bit patterns: unused slots are NaNs nobody reads
	}
	choose_timing(e, o, (uint64_t)C * T * wire);
	if (peer_join_active(e, C)) e->join_epoch++;
	if ((rc = begin_score(e, C))) return rc;
	int nchunk = 0;
	for (uint64_t t0 = 0; t0 < T; t0 += Tc, nchunk++) {
		const int b = nchunk & 1;
		const uint64_t n = T - t0 < Tc ? T - t0 : Tc;
		uint8_t *w = e->win[b].p;
		unsigned long long *wq = (unsigned long long *)w;
		float *wd = (float *)(w + (uint64_t)C * Tc * 8);
		unsigned short *wl = (unsigned short *)(w + (uint64_t)C * Tc * 12);
		uint32_t *wpc = (uint32_t *)(w + (uint64_t)C * Tc * 14);
		unsigned short *wpd = (unsigned short *)(w + (uint64_t)C * Tc * 18);
		if (nchunk >= 2) CK(cudaStreamWaitEvent(e->copy_stream, e->win_free[b], 0));
		if (packed) {
			if ((rc = copy_rows_h2d(e, wpc, Tc * 4, pc + t0, T * 4, n * 4, C, e->copy_stream))) return rc;
			if ((rc = copy_rows_h2d(e, wpd, Tc * 2, pd + t0, T * 2, n * 2, C, e->copy_stream))) return rc;
		} else {
			if ((rc = copy_rows_h2d(e, wq, Tc * 8, q8 + t0, T * 8, n * 8, C, e->copy_stream))) return rc;
			if ((rc = copy_rows_h2d(e, wd, Tc * 4, d2 + t0, T * 4, n * 4, C, e->copy_stream))) return rc;
			if ((rc = copy_rows_h2d(e, wl, Tc * 2, li + t0, T * 2, n * 2, C, e->copy_stream))) return rc;
		}
		CK(cudaEventRecord(e->win_ready[b], e->copy_stream));
		CK(cudaStreamWaitEvent(e->stream, e->win_ready[b], 0));
		if (packed) {
			CK(apo::run_unpack_p(wpc, wpd, Tc, C, n, e->d_d2stream.p, wq, wd, wl, Tc, e->stream));
			e->timing.launches++;
		}
		apo::KqParams Q{};
		Q.q8 = wq; Q.d2 = wd; Q.li = wl; Q.pitch_evals = Tc; Q.C = C; Q.T = n;
		Q.acc = e->acc.p; Q.lut = e->d_lut.p; Q.ptab = e->d_ptab.p; Q.w2 = e->W.w[2];
		Q.pair = e->d_pair.p; Q.cbf = e->d_cbf.p; Q.n0 = e->q_n0; Q.n1 = e->q_n1;
		{
			static const int dim_of[8] = {0, 1, 3, 4, 5, 6, 7, 8};
			for (int j = 0; j < 8; j++) Q.wq[j] = e->W.w[dim_of[j]];
		}
		if ((rc = record_k1_event(e, 0))) return rc;
		CK(apo::run_reward9q(Q, qv, (o->flags & APO_SCORE_RECIP) != 0, e->sm_count, e->stream));
		if ((rc = record_k1_event(e, 1))) return rc;
		e->timing.launches++;
		CK(cudaEventRecord(e->win_free[b], e->stream));
	}
	return finish_score(e, o, C, scores, counts, topk, report);
}
}  // namespace

extern "C" int apo_score_host_compact(apo_engine *e, const apo_score_opts *o, const uint64_t *q8, const float *d2, const uint16_t *li,
                                      const uint32_t *codebook, uint32_t C, uint64_t T, double *scores, uint64_t *counts,
                                      int32_t *topk, apo_corpus_report *report) {
	if (!e) return APO_E_ARG;
	const int rc = score_host_compact_impl(e, o, q8, d2, li, nullptr, nullptr, nullptr, codebook, C, T, scores, counts, topk, report);
	if (rc) { cudaStreamSynchronize(e->copy_stream); cudaStreamSynchronize(e->stream); }
	return rc;
}

extern "C" int apo_score_host_packed(apo_engine *e, const apo_score_opts *o, const uint32_t *pc, const uint16_t *pd, const uint32_t *codebook,
                                     const uint32_t *d2book, uint32_t C, uint64_t T, double *scores, uint64_t *counts, int32_t *topk,
                                     apo_corpus_report *report) {
	if (!e) return APO_E_ARG;
	if (!pc) return fail(e, APO_E_ARG, "input is NULL");
	const int rc = score_host_compact_impl(e, o, nullptr, nullptr, nullptr, pc, pd, d2book, codebook, C, T, scores, counts, topk, report);
	if (rc) { cudaStreamSynchronize(e->copy_stream); cudaStreamSynchronize(e->stream); }
	return rc;
}

// =============================================================================== Form T (dictionary indices, 3 B / evaluation)
namespace {
int tuples_args(apo_engine *e, const uint16_t *tl, const uint8_t *th, const uint32_t *tb_pc, const uint16_t *tb_pd, uint32_t n,
                const uint32_t *codebook, const uint32_t *d2book, uint32_t C, uint64_t T) {
	if (C == 0) return fail(e, APO_E_ARG, "C == 0");
	if (!codebook || !d2book || (n && (!tb_pc || !tb_pd)) || (T && (!tl || !th))) return fail(e, APO_E_ARG, "input is NULL");
	if (n > APO_TUPLES_MAX) return fail(e, APO_E_ARG, "n_tuples=%u exceeds %u", n, APO_TUPLES_MAX);
	if (T && n == 0) return fail(e, APO_E_ARG, "evaluations without a dictionary");
	return APO_OK;
}

// Streams the two index planes [C][T] through two device windows, H2D of chunk i+1 overlapped with K1t on chunk i.
int score_host_tuples_impl(apo_engine *e, const apo_score_opts *o, const uint16_t *tl, const uint8_t *th, const uint32_t *tb_pc,
                           const uint16_t *tb_pd, uint32_t n_tuples, const uint32_t *codebook, const uint32_t *d2book, uint32_t C, uint64_t T,
                           double *scores, uint64_t *counts, int32_t *topk, apo_corpus_report *report) {
	if (!o) return fail(e, APO_E_ARG, "opts is NULL");
	int rc;
	if ((rc = tuples_args(e, tl, th, tb_pc, tb_pd, n_tuples, codebook, d2book, C, T))) return rc;
	if (o->K > C) return fail(e, APO_E_ARG, "K=%u exceeds the number of candidates C=%u", o->K, C);
	if (o->K > kMaxK) return fail(e, APO_E_ARG, "K=%u exceeds the supported beam width %u", o->K, kMaxK);
	if (o->first || o->count) return fail(e, APO_E_ARG, "windows are not supported by the host-streaming calls");
	CK(cudaSetDevice(e->device));
	if ((rc = ensure_scratch(e, C, o->K))) return rc;
	const int tile = apo::kt_tile_evals();
	uint64_t Tc = (256ull << 20) / ((uint64_t)C * 3);
	Tc = Tc / tile * tile;
	if (Tc < (uint64_t)tile) Tc = tile;
	if (Tc > round_up(T ? T : 1, tile)) Tc = round_up(T ? T : 1, tile);
	for (int i = 0; i < 2; i++) CK(e->win[i].reserve((uint64_t)C * Tc * 3));
	CK(e->tups_pc.reserve(n_tuples ? n_tuples : 1));
	CK(e->tups_pd.reserve(n_tuples ? n_tuples : 1));
	CK(e->tups_d2.reserve(4096));
	choose_timing(e, o, (uint64_t)C * T * 3);
	if (peer_join_active(e, C)) e->join_epoch++;
	if ((rc = begin_score(e, C))) return rc;
	if (n_tuples) {
		CK(cudaMemcpyAsync(e->tups_pc.p, tb_pc, (uint64_t)n_tuples * 4, cudaMemcpyHostToDevice, e->stream));
		CK(cudaMemcpyAsync(e->tups_pd.p, tb_pd, (uint64_t)n_tuples * 2, cudaMemcpyHostToDevice, e->stream));
	}
	CK(cudaMemcpyAsync(e->tups_d2.p, d2book, 4096 * 4, cudaMemcpyHostToDevice, e->stream));   // bit patterns: unused slots are NaNs, flagged if an entry names one
	if ((rc = tuples_values(e, codebook, e->tups_pc.p, e->tups_pd.p, n_tuples, e->tups_d2.p, (o->flags & APO_SCORE_RECIP) != 0))) return rc;
	int nchunk = 0;
	for (uint64_t t0 = 0; t0 < T; t0 += Tc, nchunk++) {
		const int b = nchunk & 1;
		const uint64_t n = T - t0 < Tc ? T - t0 : Tc;
		unsigned short *wl = (unsigned short *)e->win[b].p;
		unsigned char *wh = e->win[b].p + (uint64_t)C * Tc * 2;
		if (nchunk >= 2) CK(cudaStreamWaitEvent(e->copy_stream, e->win_free[b], 0));
		if ((rc = copy_rows_h2d(e, wl, Tc * 2, tl + t0, T * 2, n * 2, C, e->copy_stream))) return rc;
		if ((rc = copy_rows_h2d(e, wh, Tc, th + t0, T, n, C, e->copy_stream))) return rc;
		CK(cudaEventRecord(e->win_ready[b], e->copy_stream));
		CK(cudaStreamWaitEvent(e->stream, e->win_ready[b], 0));
		apo::KtParams P{};
		P.tl = wl; P.th = wh; P.pitch_evals = Tc; P.C = C; P.T = n; P.acc = e->acc.p;
		P.tval = e->tup_val.p; P.n_tuples = n_tuples; P.bad = e->tup_bad.p;
		if ((rc = record_k1_event(e, 0))) return rc;
		CK(apo::run_reward9t(P, e->sm_count, e->stream));
		if ((rc = record_k1_event(e, 1))) return rc;
		e->timing.launches++;
		CK(cudaEventRecord(e->win_free[b], e->stream));
	}
	return finish_score(e, o, C, scores, counts, topk, report);
}
}  // namespace

extern "C" int apo_score_host_tuples(apo_engine *e, const apo_score_opts *o, const uint16_t *tl, const uint8_t *th, const uint32_t *tbook_pc,
                                     const uint16_t *tbook_pd, uint32_t n_tuples, const uint32_t *codebook, const uint32_t *d2book,
                                     uint32_t C, uint64_t T, double *scores, uint64_t *counts, int32_t *topk, apo_corpus_report *report) {
	if (!e) return APO_E_ARG;
	const int rc = score_host_tuples_impl(e, o, tl, th, tbook_pc, tbook_pd, n_tuples, codebook, d2book, C, T, scores, counts, topk, report);
	if (rc) { cudaStreamSynchronize(e->copy_stream); cudaStreamSynchronize(e->stream); e->tup_used = false; }
	return rc;
}

extern "C" int apo_tuples_upload(apo_engine *e, const uint16_t *tl, const uint8_t *th, const uint32_t *tbook_pc, const uint16_t *tbook_pd,
                                 uint32_t n_tuples, const uint32_t *codebook, const uint32_t *d2book, uint32_t C, uint64_t T) {
	if (!e) return APO_E_ARG;
	int rc;
	if ((rc = tuples_args(e, tl, th, tbook_pc, tbook_pd, n_tuples, codebook, d2book, C, T))) return rc;
	CK(cudaSetDevice(e->device));
	e->tup_C = 0; e->tup_T = 0; e->tup_n = 0;
	const uint64_t pitch = round_up(T ? T : 1, 64);                // rows start 16-byte aligned in both planes
	CK(e->tup_l.reserve((uint64_t)C * pitch));
	CK(e->tup_h.reserve((uint64_t)C * pitch));
	CK(e->tup_pc.reserve(n_tuples ? n_tuples : 1));
	CK(e->tup_pd.reserve(n_tuples ? n_tuples : 1));
	CK(e->tup_d2.reserve(4096));
	if ((rc = h2d_rows(e, e->tup_l.p, pitch * 2, tl, T * 2, T * 2, C, e->stream))) { cudaStreamSynchronize(e->stream); return rc; }
	if ((rc = h2d_rows(e, e->tup_h.p, pitch, th, T, T, C, e->stream))) { cudaStreamSynchronize(e->stream); return rc; }
	if (n_tuples) {
		CK(cudaMemcpyAsync(e->tup_pc.p, tbook_pc, (uint64_t)n_tuples * 4, cudaMemcpyHostToDevice, e->stream));
		CK(cudaMemcpyAsync(e->tup_pd.p, tbook_pd, (uint64_t)n_tuples * 2, cudaMemcpyHostToDevice, e->stream));
	}
	CK(cudaMemcpyAsync(e->tup_d2.p, d2book, 4096 * 4, cudaMemcpyHostToDevice, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	memcpy(e->tup_book, codebook, sizeof e->tup_book);
	e->tup_C = C; e->tup_T = T; e->tup_pitch = pitch; e->tup_n = n_tuples;
	return APO_OK;
}

extern "C" int apo_host_alloc(uint64_t bytes, void **out) {
	if (!out) return APO_E_ARG;
	*out = nullptr;
	const cudaError_t c = cudaMallocHost(out, bytes ? bytes : 1);
	if (c != cudaSuccess) { cudaGetLastError(); return fail(nullptr, c == cudaErrorMemoryAllocation ? APO_E_NOMEM : APO_E_CUDA, "cudaMallocHost(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(c)); }
	return APO_OK;
}

extern "C" int apo_host_free(void *p) {
	if (!p) return APO_OK;
	return cudaFreeHost(p) == cudaSuccess ? APO_OK : APO_E_CUDA;
}

extern "C" int apo_score_host(apo_engine *e, const apo_score_opts *o, const float *dims, uint32_t C, uint64_t T,
                              double *scores, uint64_t *counts, int32_t *topk, apo_corpus_report *report) {
	return score_host_rows(e, o, (const uint8_t *)dims, 36, C, T, scores, counts, topk, report);
}

extern "C" int apo_score_host_records(apo_engine *e, const apo_score_opts *o, const void *recs, uint32_t row_bytes, uint32_t C, uint64_t T,
                                      double *scores, uint64_t *counts, int32_t *topk, apo_corpus_report *report) {
	if (row_bytes != 32 && row_bytes != 16) return fail(e, APO_E_ARG, "row_bytes must be 32 (apo_record) or 16 (apo_record16)");
	return score_host_rows(e, o, (const uint8_t *)recs, row_bytes, C, T, scores, counts, topk, report);
}

extern "C" int apo_last_timing(const apo_engine *e, apo_timing *out) {
	if (!e || !out) return APO_E_ARG;
	*out = e->timing;
	return APO_OK;
}

extern "C" int apo_debug_partials(apo_engine *e, int64_t *out, uint32_t C) {
	if (!e || !out) return fail(e, APO_E_ARG, "NULL argument");
	if (C != e->last_C || !e->acc.p) return fail(e, APO_E_STATE, "no scoring call with C=%u to read back", C);
	CK(cudaSetDevice(e->device));
	CK(cudaMemcpyAsync(out, e->partials_src ? e->partials_src : e->acc.p, 8ull * ACC_PER_CAND * C, cudaMemcpyDeviceToHost, e->stream));
	CK(cudaStreamSynchronize(e->stream));
	return APO_OK;
}

// =============================================================================== multi-GPU
extern "C" int apo_comm_unique_id(uint8_t out[APO_UNIQUE_ID_BYTES]) {
	if (!out) return fail(nullptr, APO_E_ARG, "out is NULL");
	std::string err;
	if (!nccl_load(err)) return fail(nullptr, APO_E_NCCL, "%s", err.c_str());
	NcclId id;
	const int rc = g_nccl.GetUniqueId(&id);
	if (rc != 0) return fail(nullptr, APO_E_NCCL, "ncclGetUniqueId failed (%d)", rc);
	memcpy(out, id.b, APO_UNIQUE_ID_BYTES);
	return APO_OK;
}

extern "C" int apo_comm_init(apo_engine *e, int nranks, int rank, const uint8_t id[APO_UNIQUE_ID_BYTES]) {
	if (!e || !id) return fail(e, APO_E_ARG, "NULL argument");
	if (nranks < 1 || rank < 0 || rank >= nranks) return fail(e, APO_E_ARG, "bad rank %d of %d", rank, nranks);
	CK(cudaSetDevice(e->device));
	peer_teardown(e);
	if (e->comm) { g_nccl.CommDestroy(e->comm); e->comm = nullptr; }
	e->nranks = 1; e->rank = 0;
	if (nranks == 1) return APO_OK;
	std::string err;
	if (!nccl_load(err)) return fail(e, APO_E_NCCL, "%s", err.c_str());
	NcclId nid; memcpy(nid.b, id, APO_UNIQUE_ID_BYTES);
	void *comm = nullptr;
	const int rc = g_nccl.CommInitRank(&comm, nranks, nid, rank);
	if (rc != 0) return fail(e, APO_E_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
	e->comm = comm; e->nranks = nranks; e->rank = rank; e->acc_clean_words = 0; e->partials_src = nullptr;
	// one tiny allreduce now so that the first scoring call does not pay NCCL's lazy connection setup (~1 s)
	CK(e->misc.reserve(32));
	CK(cudaMemsetAsync(e->misc.p + 24, 0, 8, e->stream));
	const int wrc = g_nccl.AllReduce(e->misc.p + 24, e->misc.p + 24, 1, kNcclInt64, kNcclSum, e->comm, e->stream);
	if (wrc != 0) return fail(e, APO_E_NCCL, "ncclAllReduce (warm-up): %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(wrc) : "error");
	CK(cudaStreamSynchronize(e->stream));
	// the join itself runs over NVLink peer memory inside the scoring launch; NCCL stays as the bootstrap and the fallback
	return peer_setup(e);
}

extern "C" int apo_comm_join_mode(const apo_engine *e) {
	if (!e || e->nranks <= 1) return 0;
	return (e->peer_ok && !e->env_nccl_join) ? 2 : 1;
}

extern "C" int apo_comm_destroy(apo_engine *e) {
	if (!e) return APO_E_ARG;
	cudaSetDevice(e->device);
	if (e->own_stream) cudaStreamSynchronize(e->stream);
	peer_teardown(e);
	if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
	e->comm = nullptr; e->nranks = 1; e->rank = 0;
	return APO_OK;
}
