// CFR wave kernels (sm_100a): regret matching, strategy averaging, reach / counterfactual-value traversal
// for thousands of concurrent Liar's Dice subgames.  Replaces, per iteration and per subgame, the CPU loops
// of the reference's CFR::step (subgame_solving.cc:577-664), update_regrets (:538-575),
// compute_reach_probabilities (:54-78), write_query_to (:104-123), query_value_net scaling (:253-269) and
// compute_expected_terminal_values / compute_win_probability (:80-98, :765-789).
//
// Execution model: one thread GROUP (a warp for depth-limited subgames, a whole CTA for full-depth trees)
// owns one subgame; the group walks the tree template level by level, lanes strided over (node, hand) items.
// Per-(node,hand,action) tables live in HBM as compact fp32 [edge][hand] arrays (edge = child node - 1) so a
// group streams its subgame's tables with unit-stride, coalesced accesses; reach probabilities and node
// values are group-private scratch (shared memory for warps, global for CTA groups).
//
// One launch = backward half of iteration i-1 (consume leaf values -> regrets, regret matching, discounting,
// average-strategy accumulation) fused with the forward half of iteration i (reach -> value-net query rows,
// scalers, terminal payoffs).  The value-net kernel runs between two launches.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cfrb {

struct TemplateDev {
  int node_off, level_off, pleaf_off, term_off;
  int N, L, T, levels;
};

struct CfrDev {
  // game
  int A, H, F, Q, Qpad, Hout;
  // templates (read-only)
  const TemplateDev* tmpl;
  const int* child_begin; const int* nchild; const int* last_bid; const int* kind; const int* slot;
  const int* level_begin; const int* pleaf_node; const int* term_node;
  const unsigned char* matches;   // [H][F] num_matches(hand, face), liars_dice.h:83-91
  // wave
  const int* wave_n;              // [1] number of live subgames
  const int* sg_tmpl; const int* sg_player; const int* sg_row_off; const int* sg_act_iter;
  const float* beliefs;           // [K][2][H]
  float* mu;                      // [K][2][H] root_values_means
  int* steps;                     // [K][2]
  float* R; float* Sg; float* S; float* Snap;   // [K][table_stride]
  int table_stride;
  float* vterm; int vterm_stride; // [K][Tmax*H] terminal payoffs of the current iteration
  float* X;                       // [rows][Qpad] fp32 query rows (SIMT net) -- or nullptr
  void* Xh;                       // fp16 query tiles in UMMA core-matrix order (tensor-core net) -- or nullptr
  const float* net_out;           // [rows][Hout] raw net outputs
  float* scaler;                  // [rows] sum of opponent reach at the pseudo-leaf
  float* scratch; size_t scratch_stride;   // global scratch (CTA groups), floats per subgame
  // params
  int linear, dcfr; float dcfr_alpha, dcfr_beta, dcfr_gamma;
  int use_net;
};

template <int G>
__device__ __forceinline__ void group_sync() {
  if (G == 32) __syncwarp(); else __syncthreads();
}

// Top-down reach of both players under Sg (subgame_solving.cc:54-78), level by level.
template <int G>
__device__ __forceinline__ void reach_pass(const CfrDev& p, const TemplateDev& t, int rp, const float* __restrict__ Sg,
                                           const float* __restrict__ b, float* reach0, float* reach1, int lane) {
  const int H = p.H;
  for (int h = lane; h < H; h += G) { reach0[h] = b[h]; reach1[h] = b[H + h]; }
  group_sync<G>();
  for (int d = 0; d + 1 < t.levels; ++d) {
    const int nb = p.level_begin[t.level_off + d], ne = p.level_begin[t.level_off + d + 1];
    const int actor = rp ^ (d & 1);
    float* ra = actor ? reach1 : reach0;   // acting player's reach gets multiplied
    float* ro = actor ? reach0 : reach1;   // the other player's reach is copied
    for (int it = lane; it < (ne - nb) * H; it += G) {
      const int n = nb + it / H, h = it % H;
      const int nc = p.nchild[t.node_off + n];
      if (!nc) continue;
      const int c0 = p.child_begin[t.node_off + n];
      const float a = ra[n * H + h], o = ro[n * H + h];
      for (int j = 0; j < nc; ++j) {
        const int c = c0 + j;
        ra[c * H + h] = a * Sg[(c - 1) * H + h];
        ro[c * H + h] = o;
      }
    }
    group_sync<G>();
  }
}

// Forward half of iteration `iter`: reach, query rows + scalers for pseudo-leaves, payoffs for terminals.
template <int G>
__device__ void cfr_forward(const CfrDev& p, int k, int trav, float* reach0, float* reach1, float* lsum, int lane) {
  const TemplateDev t = p.tmpl[p.sg_tmpl[k]];
  const int H = p.H, A = p.A;
  const int rp = p.sg_player[k];
  const float* Sg = p.Sg + (size_t)k * p.table_stride;
  const float* b = p.beliefs + (size_t)k * 2 * H;
  reach_pass<G>(p, t, rp, Sg, b, reach0, reach1, lane);

  // ---- pseudo-leaves: normalisation sums + scaler (subgame_solving.cc:257-265)
  const int row0 = p.sg_row_off[k];
  for (int r = lane; r < t.L; r += G) {
    const int n = p.pleaf_node[t.pleaf_off + r];
    float s0 = 0.f, s1 = 0.f;
    for (int h = 0; h < H; ++h) { s0 += reach0[n * H + h]; s1 += reach1[n * H + h]; }
    lsum[2 * r] = s0; lsum[2 * r + 1] = s1;
    p.scaler[row0 + r] = trav == 0 ? s1 : s0;
  }
  group_sync<G>();
  // ---- query rows (write_query_to, subgame_solving.cc:104-123).  eps = 1e-80 of the reference underflows in
  // fp32; its only effect in fp64 is "all-zero reach -> uniform", which is reproduced explicitly.
  if (p.X != nullptr || p.Xh != nullptr) {
    const int Qp = p.Qpad;
    for (int it = lane; it < t.L * Qp; it += G) {
      const int r = it / Qp, q = it % Qp;
      const int n = p.pleaf_node[t.pleaf_off + r];
      const int nd = t.node_off + n;
      float v = 0.f;
      if (q == 0) {
        // depth parity of a pseudo-leaf: all pseudo-leaves sit on the last level
        v = (float)(rp ^ ((t.levels - 1) & 1));
      } else if (q == 1) {
        v = (float)trav;
      } else if (q < 2 + A) {
        v = (q - 2 == p.last_bid[nd]) ? 1.f : 0.f;
      } else if (q < 2 + A + H) {
        const float s = lsum[2 * r];
        v = s > 0.f ? reach0[n * H + (q - 2 - A)] / s : 1.f / H;
      } else if (q < 2 + A + 2 * H) {
        const float s = lsum[2 * r + 1];
        v = s > 0.f ? reach1[n * H + (q - 2 - A - H)] / s : 1.f / H;
      }
      if (p.X != nullptr) {
        p.X[(size_t)(row0 + r) * Qp + q] = v;
      } else {
        // fp16 tile in UMMA K-major core-matrix order (leaf_mlp_tc.cuh umma_kmajor_offset_halves, R = 128)
        const int R = row0 + r, rr = R & 127;
        reinterpret_cast<__half*>(p.Xh)[(size_t)(R >> 7) * 128 * Qp + (q >> 3) * 1024 + (rr >> 3) * 64 + (rr & 7) * 8 + (q & 7)] =
            __float2half_rn(v);
      }
    }
  }
  // ---- terminals (compute_expected_terminal_values, subgame_solving.cc:80-98; win probability :765-789)
  // term_node holds three lists of length T: node id, challenged bid (= parent's last_bid, :287), node depth.
  float* vt = p.vterm + (size_t)k * p.vterm_stride;
  const float* ropp = trav == 0 ? reach1 : reach0;
  for (int it = lane; it < t.T * H; it += G) {
    const int z = it / H, h = it % H;
    const int n = p.term_node[t.term_off + z];
    const int pbid = p.term_node[t.term_off + t.T + z];
    const int ndepth = p.term_node[t.term_off + 2 * t.T + z];
    const int quantity = 1 + pbid / p.F, face = pbid % p.F;   // unpack_action, liars_dice.h:74-80
    const float* ro = ropp + n * H;
    // P(h) = sum of opponent reach over hands g with matches(g) >= quantity - matches(h): the suffix sum of
    // the match-count histogram the reference builds (:770-779), evaluated directly.
    const int need = quantity - (int)p.matches[h * p.F + face];
    float win = 0.f, tot = 0.f;
    for (int g = 0; g < H; ++g) {
      const float r = ro[g];
      tot += r;
      if ((int)p.matches[g * p.F + face] >= need) win += r;
    }
    const float v = win * 2.f - tot;
    // state.player_id of a terminal = the bidder; payoff is negated iff that is not the traverser (:290)
    const int pl = rp ^ (ndepth & 1);
    vt[z * H + h] = (pl != trav) ? -v : v;
  }
}

// Backward half of iteration with traverser `trav` (update_regrets :538-575 and step :577-664).
template <int G>
__device__ void cfr_backward(const CfrDev& p, int k, int trav, float* reach0, float* reach1, float* val, int lane) {
  const TemplateDev t = p.tmpl[p.sg_tmpl[k]];
  const int H = p.H;
  const int rp = p.sg_player[k];
  float* R = p.R + (size_t)k * p.table_stride;
  float* Sg = p.Sg + (size_t)k * p.table_stride;
  float* S = p.S + (size_t)k * p.table_stride;
  const int row0 = p.sg_row_off[k];
  // leaf values = net(query) * scaler (subgame_solving.cc:266-282); terminals from the forward half
  for (int it = lane; it < t.L * H; it += G) {
    const int r = it / H, h = it % H;
    const int n = p.pleaf_node[t.pleaf_off + r];
    val[n * H + h] = p.use_net ? p.net_out[(size_t)(row0 + r) * p.Hout + h] * p.scaler[row0 + r] : 0.f;
  }
  const float* vt = p.vterm + (size_t)k * p.vterm_stride;
  for (int it = lane; it < t.T * H; it += G) {
    const int z = it / H, h = it % H;
    val[p.term_node[t.term_off + z] * H + h] = vt[z * H + h];
  }
  group_sync<G>();
  // bottom-up (reverse BFS order == decreasing level)
  for (int d = t.levels - 2; d >= 0; --d) {
    const int nb = p.level_begin[t.level_off + d], ne = p.level_begin[t.level_off + d + 1];
    const bool mine = (rp ^ (d & 1)) == trav;
    for (int it = lane; it < (ne - nb) * H; it += G) {
      const int n = nb + it / H, h = it % H;
      const int nc = p.nchild[t.node_off + n];
      if (!nc) continue;
      const int c0 = p.child_begin[t.node_off + n];
      float v = 0.f;
      if (mine) {
        for (int j = 0; j < nc; ++j) v += val[(c0 + j) * H + h] * Sg[(c0 + j - 1) * H + h];
        for (int j = 0; j < nc; ++j) {
          const int e = (c0 + j - 1) * H + h;
          R[e] = (R[e] + val[(c0 + j) * H + h]) - v;
        }
      } else {
        for (int j = 0; j < nc; ++j) v += val[(c0 + j) * H + h];
      }
      val[n * H + h] = v;
    }
    group_sync<G>();
  }
  // root value running mean (:579-590) and discounts (:592-617)
  const int s = p.steps[2 * k + trav];
  {
    const float alpha = p.linear ? 2.f / (s + 2) : 1.f / (s + 1);
    float* mu = p.mu + ((size_t)k * 2 + trav) * H;
    for (int h = lane; h < H; h += G) mu[h] += (val[h] - mu[h]) * alpha;
  }
  float pos = 1.f, neg = 1.f, strat = 1.f;
  {
    const float ns = (float)(s + 1);
    if (p.linear) {
      pos = neg = strat = ns / (ns + 1.f);
    } else if (p.dcfr) {
      pos = p.dcfr_alpha >= 5.f ? 1.f : powf(ns, p.dcfr_alpha) / (powf(ns, p.dcfr_alpha) + 1.f);
      neg = p.dcfr_beta <= -5.f ? 0.f : powf(ns, p.dcfr_beta) / (powf(ns, p.dcfr_beta) + 1.f);
      strat = powf(ns / (ns + 1.f), p.dcfr_gamma);
    }
  }
  // regret matching (:619-634), traverser reach under the new strategy (:636-638), regret discount,
  // sum-strategy update (:639-661); top-down so reach of the parent is ready.
  float* rt = trav ? reach1 : reach0;
  const float* b = p.beliefs + ((size_t)k * 2 + trav) * H;
  for (int h = lane; h < H; h += G) rt[h] = b[h];
  group_sync<G>();
  for (int d = 0; d + 1 < t.levels; ++d) {
    const int nb = p.level_begin[t.level_off + d], ne = p.level_begin[t.level_off + d + 1];
    const bool mine = (rp ^ (d & 1)) == trav;
    for (int it = lane; it < (ne - nb) * H; it += G) {
      const int n = nb + it / H, h = it % H;
      const int nc = p.nchild[t.node_off + n];
      if (!nc) continue;
      const int c0 = p.child_begin[t.node_off + n];
      const float rn = rt[n * H + h];
      if (mine) {
        float sum = 0.f;
        for (int j = 0; j < nc; ++j) sum += fmaxf(R[(c0 + j - 1) * H + h], 0.f);
        const float inv = sum > 0.f ? 1.f / sum : 0.f, uni = 1.f / nc;
        for (int j = 0; j < nc; ++j) {
          const int e = (c0 + j - 1) * H + h;
          const float r = R[e];
          const float sg = sum > 0.f ? fmaxf(r, 0.f) * inv : uni;
          Sg[e] = sg;
          R[e] = r * (r > 0.f ? pos : neg);
          S[e] = S[e] * strat + rn * sg;
          rt[(c0 + j) * H + h] = rn * sg;
        }
      } else {
        for (int j = 0; j < nc; ++j) rt[(c0 + j) * H + h] = rn;
      }
    }
    group_sync<G>();
  }
  if (lane == 0) p.steps[2 * k + trav] = s + 1;
}

// iter: global iteration index of the forward half.  do_b: run backward half of iteration iter-1 first.
template <int G>
__global__ void __launch_bounds__(256) cfr_iter_kernel(CfrDev p, int iter, int do_b, int do_f, int smem_floats_per_group) {
  extern __shared__ float smem[];
  const int groups_per_cta = blockDim.x / G;
  const int gid = threadIdx.x / G, lane = threadIdx.x % G;
  const int k = blockIdx.x * groups_per_cta + gid;
  const int n = *p.wave_n;
  if (k >= n) return;   // uniform per group (and per CTA when G == blockDim.x)
  const TemplateDev t = p.tmpl[p.sg_tmpl[k]];
  const int NH = t.N * p.H;
  float* base = (p.scratch != nullptr) ? p.scratch + (size_t)k * p.scratch_stride : smem + (size_t)gid * smem_floats_per_group;
  float* reach0 = base; float* reach1 = base + NH; float* val = base + 2 * NH; float* lsum = base + 3 * NH;
  if (do_b) {
    cfr_backward<G>(p, k, (iter - 1) & 1, reach0, reach1, val, lane);
    group_sync<G>();
  }
  // sampling-strategy snapshot for RlRunner (recursive_solving.cc:168-174): state after `iter` iterations
  if (p.sg_act_iter[k] == iter) {
    const float* Sg = p.Sg + (size_t)k * p.table_stride;
    float* Sn = p.Snap + (size_t)k * p.table_stride;
    for (int i = lane; i < (t.N - 1) * p.H; i += G) Sn[i] = Sg[i];
  }
  if (do_f) cfr_forward<G>(p, k, iter & 1, reach0, reach1, lsum, lane);
}

// Wave initialisation == CFR constructor (subgame_solving.cc:509-524): uniform last strategy, zero regrets,
// sum = uniform * reach-under-uniform of the acting player (get_uniform_reach_weigted_strategy :125-149).
template <int G>
__global__ void __launch_bounds__(256) cfr_init_kernel(CfrDev p, int smem_floats_per_group) {
  extern __shared__ float smem[];
  const int groups_per_cta = blockDim.x / G;
  const int gid = threadIdx.x / G, lane = threadIdx.x % G;
  const int k = blockIdx.x * groups_per_cta + gid;
  if (k >= *p.wave_n) return;
  const TemplateDev t = p.tmpl[p.sg_tmpl[k]];
  const int H = p.H, NH = t.N * H;
  float* base = (p.scratch != nullptr) ? p.scratch + (size_t)k * p.scratch_stride : smem + (size_t)gid * smem_floats_per_group;
  float* reach0 = base; float* reach1 = base + NH;
  float* R = p.R + (size_t)k * p.table_stride;
  float* Sg = p.Sg + (size_t)k * p.table_stride;
  float* S = p.S + (size_t)k * p.table_stride;
  const int rp = p.sg_player[k];
  for (int n = 0; n < t.N; ++n) {
    const int nc = p.nchild[t.node_off + n];
    if (!nc) continue;
    const int c0 = p.child_begin[t.node_off + n];
    const float u = 1.f / nc;
    for (int it = lane; it < nc * H; it += G) { Sg[(c0 - 1) * H + it] = u; R[(c0 - 1) * H + it] = 0.f; }
  }
  group_sync<G>();
  reach_pass<G>(p, t, rp, Sg, p.beliefs + (size_t)k * 2 * H, reach0, reach1, lane);
  for (int d = 0; d + 1 < t.levels; ++d) {
    const int nb = p.level_begin[t.level_off + d], ne = p.level_begin[t.level_off + d + 1];
    const float* ra = (rp ^ (d & 1)) ? reach1 : reach0;
    for (int it = lane; it < (ne - nb) * H; it += G) {
      const int n = nb + it / H, h = it % H;
      const int nc = p.nchild[t.node_off + n];
      if (!nc) continue;
      const int c0 = p.child_begin[t.node_off + n];
      for (int j = 0; j < nc; ++j) S[(c0 + j - 1) * H + h] = Sg[(c0 + j - 1) * H + h] * ra[n * H + h];
    }
  }
  for (int i = lane; i < 2 * H; i += G) p.mu[(size_t)k * 2 * H + i] = 0.f;
  if (lane < 2) p.steps[2 * k + lane] = 0;
}

}  // namespace cfrb
