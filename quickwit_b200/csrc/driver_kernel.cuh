// driver_kernel.cuh — posting-driven evaluation of "one term AND column filters" queries.
//
// The window engine (kernels.cuh) pays per doc-id window: bitmaps to clear, a program to interpret, a block-wide
// barrier per clause, a sweep over the window at collect time. A query whose matches all come from ONE posting
// list — a single term (BASELINE config 1), or a required term narrowed by fast-field range / exists filters
// (config 3: term AND timestamp range, top-K by timestamp) — touches a few percent of the docs, and the window
// overheads dominate. This kernel walks the term's posting blocks instead, one warp per 128-posting block:
//   decode doc ids (+ tfs)  ->  per posting: probe the filter columns  ->  count the hit
//   MODE_HIST:    11-bit digit of the composite sort key -> the split's radix histogram
//   MODE_COLLECT: cheap 11-bit pre-filter, then the full key against the threshold -> candidate list
// Work is proportional to the postings of the driving term; there is no per-window state at all.
// Replaces, for these shapes, the same tantivy loop as the window engine: TermScorer / Intersection with
// FastFieldRangeQuery docsets + QuickwitSegmentCollector::collect (quickwit-search/src/leaf.rs:637,
// collector.rs:373-470). Same composite key, same threshold / radix select / k_select as the other paths.
#pragma once
#include "kernels.cuh"

namespace qwk {

#define QD_WARPS 8
#define QD_MAX_FILTERS 4

struct DrvParams {
  const DSplitPlan* plans;
  const DInstr* instrs;
  const DCol* cols;
  const DThresh* thresh;
  const uint32_t* first_work;  // prefix over splits of the driving term's block counts; [n_splits + 1]
  uint32_t n_splits, total_work;
  uint32_t level, use_prefix;  // MODE_HIST
  uint32_t stride;             // threshold sample: work item k of split s = block k * stride + s % stride (1 = every block)
};

template <int MODE>
__global__ void __launch_bounds__(QD_WARPS * 32) k_driver(const DrvParams p) {
  __shared__ uint32_t s_hist[QW_HIST_BINS];
  __shared__ unsigned long long s_cnt[2];  // hits, eligible of the current split
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // this block's chunk of the flat (split, posting block) list
  const uint32_t w_begin = (uint32_t)(((uint64_t)p.total_work * blockIdx.x) / gridDim.x);
  const uint32_t w_end = (uint32_t)(((uint64_t)p.total_work * (blockIdx.x + 1)) / gridDim.x);
  if (w_begin >= w_end) return;
  uint32_t split = 0;
  {
    uint32_t a = 0, b = p.n_splits;
    while (b - a > 1) {
      const uint32_t mid = (a + b) >> 1;
      if (__ldg(p.first_work + mid) <= w_begin) a = mid; else b = mid;
    }
    split = a;
  }
  if (MODE == MODE_HIST) for (uint32_t i = tid; i < QW_HIST_BINS; i += QD_WARPS * 32) s_hist[i] = 0;
  if (tid < 2) s_cnt[tid] = 0;
  __syncthreads();

  for (uint32_t seg = w_begin; seg < w_end;) {
    while (__ldg(p.first_work + split + 1) <= seg) split++;
    const uint32_t fw = __ldg(p.first_work + split);
    const uint32_t seg_end = min(w_end, __ldg(p.first_work + split + 1));
    const DSplitPlan& P = p.plans[split];
    const DKeySpec& ks = P.key;
    const DCol* cols = p.cols + P.col_base;
    const uint8_t* base = (const uint8_t*)P.data_base;
    const DInstr* prog = p.instrs + P.instr_base;  // [BOOL_BEGIN, TERM, filter x n, BOOL_END]
    const DInstr& term = prog[1];
    const uint32_t n_filters = P.n_instr - 3;
    const DThresh& T = p.thresh[split];
    const Key thr{T.key[0], T.key[1], T.key[2]};
    const uint32_t thr_top = (uint32_t)(thr.w0 >> 53);
    const bool scored = (term.flags & IF_SCORED) != 0;
    const bool has_tf = (term.flags & IF_HAS_TF) != 0, has_fn = (term.flags & IF_HAS_FN) != 0;
    const float* tab = (const float*)P.bm25_tab[term.r];
    const QwSkip* skips = (const QwSkip*)(base + term.c);
    const uint8_t* tdata = base + term.a;
    const uint32_t sa_present = P.sa.present, max_hits = P.max_hits;
    uint32_t my_hits = 0, my_elig = 0;

    for (uint32_t w = seg + warp; w < seg_end; w += QD_WARPS) {
      const uint32_t bb = (w - fw) * p.stride + (p.stride > 1 ? split % p.stride : 0);
      const uint8_t* blk = tdata + __ldg(&skips[bb].byte_off);
      const uint4 h = __ldg(reinterpret_cast<const uint4*>(blk));  // QwSkip: last_doc, prev_last_doc, byte_off, bits/count
      const uint32_t prev = h.y, doc_bits = h.w & 0xFF, tf_bits = (h.w >> 8) & 0xFF, count = h.w >> 16;
      const uint4* dp = reinterpret_cast<const uint4*>(blk + 16);
      uint32_t v[4] = {0, 0, 0, 0};
      if (doc_bits) {
        const uint32_t bitpos = lane * doc_bits, wi = bitpos >> 5, sh = bitpos & 31;
        const uint4 A = __ldg(dp + wi);
        const uint4 B = (sh + doc_bits > 32) ? __ldg(dp + wi + 1) : make_uint4(0, 0, 0, 0);
        const uint32_t mask = 0xFFFFFFFFu >> (32 - doc_bits);
        v[0] = __funnelshift_r(A.x, B.x, sh) & mask; v[1] = __funnelshift_r(A.y, B.y, sh) & mask;
        v[2] = __funnelshift_r(A.z, B.z, sh) & mask; v[3] = __funnelshift_r(A.w, B.w, sh) & mask;
      }
      const uint32_t d0 = v[0] + 1, d1 = d0 + v[1] + 1, d2 = d1 + v[2] + 1, d3 = d2 + v[3] + 1;
      uint32_t incl = d3;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(QW_FULL, incl, o);
        if ((int)lane >= o) incl += n;
      }
      const uint32_t basev = prev + (incl - d3);  // mod 2^32
      const uint32_t doc[4] = {basev + d0, basev + d1, basev + d2, basev + d3};
      uint32_t tf[4] = {1, 1, 1, 1};
      if (scored && has_tf && tf_bits) {
        const uint4* tp = dp + doc_bits;
        const uint32_t bitpos = lane * tf_bits, wi = bitpos >> 5, sh = bitpos & 31;
        const uint4 A = __ldg(tp + wi);
        const uint4 B = (sh + tf_bits > 32) ? __ldg(tp + wi + 1) : make_uint4(0, 0, 0, 0);
        const uint32_t mask = 0xFFFFFFFFu >> (32 - tf_bits);
        tf[0] = __funnelshift_r(A.x, B.x, sh) & mask; tf[1] = __funnelshift_r(A.y, B.y, sh) & mask;
        tf[2] = __funnelshift_r(A.z, B.z, sh) & mask; tf[3] = __funnelshift_r(A.w, B.w, sh) & mask;
      }
      uint32_t fnw = 0x01010101u;  // no fieldnorms: constant fieldnorm id 1
      if (scored && has_fn) fnw = __ldg(reinterpret_cast<const uint32_t*>(blk + 16u + 16u * (doc_bits + tf_bits)) + lane);

#pragma unroll
      for (int j = 0; j < 4; j++) {
        bool on = lane * 4 + j < count;
        // ---- required column filters (FastFieldRangeQuery / ExistsQuery docsets) ------------------------------
        for (uint32_t f = 0; f < n_filters && on; f++) {
          const DInstr& fi = prog[2 + f];
          if (fi.r == 0xFFFFFFFFu) { on = false; break; }  // column absent from the split: nothing matches
          const DCol& c = cols[fi.r];
          uint64_t a, b;
          col_range(base, c, doc[j], a, b);
          if (fi.op == OP_EXISTS) on = a != b;
          else {
            bool hit = false;
            for (uint64_t k = a; k < b && !hit; k++) {
              const uint64_t mv = c.min_value + c.gcd * col_raw(base, c, k);
              hit = mv >= fi.a && mv <= fi.b;
            }
            on = hit;
          }
        }
        my_hits += on ? 1u : 0u;
        if (!max_hits) continue;
        // ---- BM25 of the posting: weight * tf-factor (Bm25Weight::score), same table / divide as fold_block -----
        float sc = 0.0f;
        if (scored && on) {
          const uint32_t fn = (fnw >> (8 * j)) & 0xFFu;
          float tfn;
          if (tf[j] < QW_TFF_ROWS) tfn = __ldg(tab + 256 + tf[j] * 256 + fn);
          else { const float tff = (float)tf[j]; tfn = __fdiv_rn(tff, __fadd_rn(tff, __ldg(tab + fn))); }
          sc = __fmul_rn(term.f, tfn);
        }
        if (MODE == MODE_COLLECT) {
          if (!on) continue;
          if (sa_present) {
            const DocKey dk = doc_key(P, ks, cols, base, doc[j], sc);
            if (dk.eligible) {
              my_elig++;
              if (key_ge(dk.key, thr)) {
                const uint32_t pos = atomicAdd((uint32_t*)P.out_cand_count, 1u);
                if (pos < QW_CAND_CAP) { uint64_t* c = (uint64_t*)P.out_cands + 3ull * pos; c[0] = dk.key.w0; c[1] = dk.key.w1; c[2] = dk.key.w2; }
              }
            }
          } else if (key_top11(P, ks, cols, base, doc[j], sc) >= thr_top) {
            const DocKey dk = doc_key(P, ks, cols, base, doc[j], sc);
            if (key_ge(dk.key, thr)) {
              const uint32_t pos = atomicAdd((uint32_t*)P.out_cand_count, 1u);
              if (pos < QW_CAND_CAP) { uint64_t* c = (uint64_t*)P.out_cands + 3ull * pos; c[0] = dk.key.w0; c[1] = dk.key.w1; c[2] = dk.key.w2; }
            }
          }
        } else {
          uint32_t digit = 0;
          bool ranked = false;
          if (on) {
            if (p.level == 0 && !sa_present) { digit = key_top11(P, ks, cols, base, doc[j], sc); ranked = true; }
            else {
              const DocKey dk = doc_key(P, ks, cols, base, doc[j], sc);
              ranked = dk.eligible && (!p.use_prefix || key_prefix_eq(dk.key, T.key, T.prefix_bits));
              digit = key_digit(dk.key, p.level);
            }
          }
          warp_count_uniform(s_hist, digit, ranked, lane);
        }
      }
    }

    // ---- end of this split's segment: counters and histogram go out ---------------------------------------------
    if (MODE == MODE_COLLECT) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        my_hits += __shfl_down_sync(QW_FULL, my_hits, o);
        my_elig += __shfl_down_sync(QW_FULL, my_elig, o);
      }
      if (lane == 0) {
        if (my_hits) atomicAdd(&s_cnt[0], (unsigned long long)my_hits);
        if (my_elig) atomicAdd(&s_cnt[1], (unsigned long long)my_elig);
      }
    }
    __syncthreads();
    if (MODE == MODE_COLLECT) {
      if (tid == 0) {
        const unsigned long long hits = s_cnt[0];
        if (hits) atomicAdd((unsigned long long*)P.out_num_hits, hits);
        const unsigned long long elig = sa_present ? s_cnt[1] : (max_hits ? hits : 0ull);  // without search_after every hit is eligible
        if (elig) atomicAdd((unsigned long long*)P.out_num_hits + 1, elig);
        s_cnt[0] = s_cnt[1] = 0;
      }
    } else {
      uint32_t* gh = (uint32_t*)P.out_hist;
      for (uint32_t i = tid; i < QW_HIST_BINS; i += QD_WARPS * 32) {
        const uint32_t c = s_hist[i];
        if (c) { atomicAdd(&gh[i], c); s_hist[i] = 0; }
      }
    }
    __syncthreads();
    seg = seg_end;
  }
}

}  // namespace qwk
