// apo_kernels.h — host-visible launch interface of apo_kernels.cu (internal, not the ABI).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/apo_b200.h"

namespace apo {

struct Weights { double w[APO_NDIM]; };

struct K1Params {
	const uint8_t *base;        // first evaluation of candidate 0 inside the window
	uint64_t pitch_bytes;       // candidate row pitch
	uint32_t C;
	uint32_t tiles_per_cand;    // filled by run_reward9
	uint64_t T;                 // evaluations per candidate in this launch
	uint64_t total_tiles;       // filled by run_reward9
	long long *acc;             // accumulator vector (see apo_device.cuh)
	const double *lut;          // [0,512) total weight per presence mask, [512,1024) its reciprocal,
	                            // [1024,1088) categorical product table (apo_device.cuh CAT_*)
	Weights W;
};

struct FinalizeParams {
	const long long *acc;
	uint32_t C, K;
	int nranks;
	int with_corpus;
	double *scores;             // [C]
	uint64_t *counts;           // [C]
	unsigned long long *keys;   // [C] scratch
	unsigned long long *sel_key;// [K] scratch
	int32_t *sel_idx;           // [K] scratch
	int32_t *topk;              // [K]
	apo_corpus_report *report;
};

struct K2Params {
	const apo_record *recs;
	uint64_t T;
	uint64_t idx_base;
	uint32_t C;
	int rank;
	long long *acc;
	unsigned long long *ex_scratch;  // [18], preset to ~0
	unsigned int *ticket;            // preset to 0
	const double *lut;
	Weights W;
	int fuse_finalize;
	FinalizeParams fin;
};

int k1_tile_evals(int row, int variant);
cudaError_t run_reward9(K1Params P, int row, int variant, bool recip, int sm_count, cudaStream_t st);
cudaError_t run_detect6(const K2Params &P, int sm_count, cudaStream_t st);
cudaError_t run_finalize(const FinalizeParams &F, cudaStream_t st);
cudaError_t run_gen_dims(float *out, uint64_t pitch_evals, uint64_t seed, uint32_t c0, uint32_t C, uint64_t t0, uint64_t T,
                         uint32_t agent_permille, cudaStream_t st);
cudaError_t run_gen_records(apo_record *out, uint64_t pitch, uint64_t seed, uint32_t stream, uint32_t c0, uint32_t C,
                            uint64_t t0, uint64_t T, uint32_t agent_permille, cudaStream_t st);
cudaError_t run_gen_records16(apo_record16 *out, uint64_t pitch, uint64_t seed, uint32_t stream, uint32_t c0, uint32_t C,
                              uint64_t t0, uint64_t T, uint32_t agent_permille, cudaStream_t st);
cudaError_t run_reward_batch(const apo_record *recs, uint64_t n, const Weights &W, const double *lut, double *dims,
                             uint32_t *masks, double *finals, cudaStream_t st);

}  // namespace apo
