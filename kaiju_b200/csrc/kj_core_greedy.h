// kj_core_greedy.h -- Greedy mode (stub while MEM is brought up)
#pragma once
#include "kj_core.h"
static KJ_HD uint32_t kj_greedy_scratch_bytes(const KjRunParams& rp) { return 64; }
static KJ_DEV uint32_t kj_classify_greedy(KjWarpCtx& cx, KjQueue& q, int n1, int n2, uint32_t& best_out) { best_out = 0; return KJ_TAX_BAD; }
