// apo_format.h — Form R <-> Form R16 (include/apo_b200.h), shared by host and device code.
#pragma once
#include <cstdint>
#include "../../include/apo_b200.h"

#ifdef __CUDACC__
#define APO_HD __host__ __device__ __forceinline__
#else
#define APO_HD inline
#endif

namespace apo {

// true when r survives the 16-byte packing with every output of the path unchanged
APO_HD bool representable16(const apo_record &r) {
	return r.toolCalls <= 65535u && r.toolFail <= 65535u && r.toolSucc + r.toolFail == r.toolCalls &&
	       r.feedback <= 2 && r.mode <= 4;
}

APO_HD apo_record16 pack16(const apo_record &r) {
	apo_record16 p;
	p.hdr = (uint16_t)((r.feedback & 3u) | ((r.flags & APO_F_ERRORS) ? 4u : 0u) | ((r.flags & APO_F_ENDED) ? 8u : 0u) |
	                   ((r.flags & APO_F_VALID) ? 16u : 0u) | ((r.flags & APO_F_FAILSPAN) ? 32u : 0u) | ((uint32_t)(r.mode & 7u) << 6));
	p.userMsgs = (uint8_t)(r.userMsgs < 255 ? r.userMsgs : 255);
	p.asstMsgs = (uint8_t)(r.asstMsgs < 255 ? r.asstMsgs : 255);
	p.toolCalls = (uint16_t)r.toolCalls;
	p.toolFail = (uint16_t)r.toolFail;
	p.llmCalls = (uint8_t)(r.llmCalls < 255u ? r.llmCalls : 255u);
	p.durClass = r.durClass;
	p.tokens = (uint16_t)(r.tokens < 65535u ? r.tokens : 65535u);
	p.toolDurMs = r.toolDurMs;
	return p;
}

APO_HD apo_record unpack16(const apo_record16 &p) {
	apo_record r;
	const uint32_t h = p.hdr;
	r.feedback = (uint8_t)(h & 3u);
	r.flags = (uint8_t)(((h & 4u) ? APO_F_ERRORS : 0u) | ((h & 8u) ? APO_F_ENDED : 0u) | ((h & 16u) ? APO_F_VALID : 0u) |
	                    ((h & 32u) ? APO_F_FAILSPAN : 0u));
	r.mode = (uint8_t)((h >> 6) & 7u);
	r.durClass = p.durClass;
	r.userMsgs = p.userMsgs;
	r.asstMsgs = p.asstMsgs;
	r.toolCalls = p.toolCalls;
	r.toolFail = p.toolFail;
	r.toolSucc = (uint32_t)p.toolCalls - (uint32_t)p.toolFail;
	r.llmCalls = p.llmCalls;
	r.tokens = p.tokens;
	r.toolDurMs = p.toolDurMs;
	return r;
}

}  // namespace apo
