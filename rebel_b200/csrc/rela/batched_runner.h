// BatchedRlRunner: self-play of K Liar's Dice games in lock-step on one GPU — the B200 counterpart of the reference's
// RlRunner (recursive_solving.h:40-86, recursive_solving.cc:160-275), which plays ONE game per CPU thread.
//
// A "wave" solves the current subgame of every game at once through the C ABI (include/cfrb200.h): cfrb_begin_wave ->
// cfrb_run(num_iters) -> snapshot of the sampling strategy at each game's act_iteration + training examples.  Everything
// the reference does per game on the host stays per game on the host, with the reference's own random-number draw order
// and fp64 belief arithmetic, so that game 0 of a runner seeded with s replays RlRunner(seed = s) exactly as long as the
// solver's strategies agree:
//   act_iteration ~ U{0..num_iters}                                   recursive_solving.cc:168-169
//   sample_state_to_leaf / sample_state_single                        :192-246 / :248-275
//       br_sampler ~ U{0,1}; per node eps ~ U[0,1) (float); random action or hand ~ beliefs, action ~ policy[hand]
//   belief update + eps-normalisation                                 :41-44, :235-245
//   two training examples per solved subgame                          subgame_solving.cc:672-676
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/cfrb200.h"
#include "params.h"

namespace rela {

// Receives blocks of examples: q [n][Q], v [n][H].  Returns false to stop the runner (e.g. replay closed).
using ExampleSink = std::function<bool(const float* q, int q_dim, const float* v, int v_dim, int n)>;

class BatchedRlRunner {
 public:
  BatchedRlRunner(const liars_dice::RecursiveSolvingParams& cfg, int device, int seed)
      : cfg_(cfg), K_(std::max(1, cfg.concurrent_games)) {
    const auto& sp = cfg.subgame_params;
    cfrb_config c{};
    c.solver = sp.use_cfr ? CFRB_SOLVER_CFR : CFRB_SOLVER_FP;   // build_solver (subgame_solving.cc:791-800)
    c.optimistic = sp.optimistic;
    c.num_dice = cfg.num_dice; c.num_faces = cfg.num_faces; c.max_depth = sp.max_depth; c.num_iters = sp.num_iters;
    c.linear_update = sp.linear_update; c.dcfr = sp.dcfr; c.dcfr_alpha = sp.dcfr_alpha; c.dcfr_beta = sp.dcfr_beta;
    c.dcfr_gamma = sp.dcfr_gamma; c.max_subgames = K_; c.device = device; c.net_mode = cfg.net_mode; c.hidden = 256;
    c.state_dtype = cfg.state_dtype;
    check(cfrb_create(&c, &h_), "cfrb_create");
    A_ = cfrb_num_actions(h_); H_ = cfrb_num_hands(h_); Q_ = cfrb_query_size(h_);
    stride_ = cfrb_table_stride(h_);
    trees_.resize(A_);   // root bids -1 .. A-2
    for (int b = -1; b <= A_ - 2; ++b) tree(b);   // built up front: the per-game walk runs on several threads and only reads them
    games_.resize(K_);
    for (int g = 0; g < K_; ++g) {
      // game 0 carries the caller's seed verbatim (parity with RlRunner(seed)); the others get decorrelated streams
      games_[g].gen.seed(g == 0 ? (uint32_t)seed : (uint32_t)(seed * 1000003u + 7919u * (uint32_t)g));
      resetGame(games_[g]);
    }
    last_bid_.resize(K_); player_.resize(K_); act_.resize(K_);
    beliefs_.resize((size_t)K_ * 2 * H_); snap_.resize((size_t)K_ * stride_);
    ex_q_.resize((size_t)K_ * 2 * Q_); ex_v_.resize((size_t)K_ * 2 * H_);
  }
  ~BatchedRlRunner() { if (h_) cfrb_destroy(h_); }
  BatchedRlRunner(const BatchedRlRunner&) = delete;
  BatchedRlRunner& operator=(const BatchedRlRunner&) = delete;

  void setWeights(const std::vector<float>& flat, uint64_t version) {
    check(cfrb_set_weights(h_, flat.data(), flat.size(), version), "cfrb_set_weights");
  }
  uint64_t weightsVersion() const { return cfrb_weights_version(h_); }
  int games() const { return K_; }
  int64_t subgamesSolved() const { return subgames_solved_; }

  // One wave: solve the current subgame of every game, emit 2 examples per subgame, advance every game.
  bool step(const ExampleSink& sink) {
    const int iters = cfg_.subgame_params.num_iters;
    for (int g = 0; g < K_; ++g) {
      Game& G = games_[g];
      G.act_iteration = std::uniform_int_distribution<>(0, iters)(G.gen);   // recursive_solving.cc:168-169
      last_bid_[g] = G.last_bid; player_[g] = G.player; act_[g] = G.act_iteration;
      std::copy(G.beliefs.begin(), G.beliefs.end(), beliefs_.begin() + (size_t)g * 2 * H_);
    }
    check(cfrb_begin_wave(h_, K_, last_bid_.data(), player_.data(), beliefs_.data(), act_.data()), "cfrb_begin_wave");
    check(cfrb_run(h_, iters, nullptr), "cfrb_run");
    check(cfrb_fetch_compact(h_, /*snapshot*/ 0, snap_.data()), "cfrb_fetch_compact");
    check(cfrb_examples(h_, ex_q_.data(), ex_v_.data()), "cfrb_examples");
    subgames_solved_ += K_;
    // every game owns its random stream and beliefs, so the walk is split over a few host threads without changing any result
    auto walk = [this](int g0, int g1) {
      for (int g = g0; g < g1; ++g) {
        Game& G = games_[g];
        const double* sigma = snap_.data() + (size_t)g * stride_;
        if (cfg_.sample_leaf) sampleToLeaf(G, sigma); else sampleSingle(G, sigma);
        if (G.last_bid == A_ - 1) resetGame(G);   // terminal: RlRunner::step returns, the next call starts a new game
      }
    };
    const int T = K_ >= 2048 ? 4 : 1;
    if (T == 1) {
      walk(0, K_);
    } else {
      std::vector<std::thread> th;
      for (int i = 1; i < T; ++i) th.emplace_back(walk, (int)((int64_t)K_ * i / T), (int)((int64_t)K_ * (i + 1) / T));
      walk(0, K_ / T);
      for (auto& x : th) x.join();
    }
    return sink(ex_q_.data(), Q_, ex_v_.data(), H_, 2 * K_);
  }

 private:
  struct Game {
    int last_bid = -1, player = 0, act_iteration = 0;
    std::vector<double> beliefs;   // [2][H]
    std::mt19937 gen;
  };

  void check(int rc, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + cfrb_last_error());
  }
  void resetGame(Game& G) {   // recursive_solving.cc:161-163
    G.last_bid = -1; G.player = 0;
    G.beliefs.assign((size_t)2 * H_, 1.0 / H_);
  }
  const std::vector<cfrb_node>& tree(int root_bid) {
    auto& t = trees_[root_bid + 1];
    if (t.empty()) {
      t.resize(cfrb_max_nodes(h_));
      int n = cfrb_tree_template(h_, root_bid, 0, t.data(), (int)t.size());
      if (n < 0) throw std::runtime_error(cfrb_last_error());
      t.resize(n);
    }
    return t;
  }
  void bidRange(int last_bid, int* lo, int* hi) const {   // liars_dice.h:110-115
    if (last_bid < 0) { *lo = 0; *hi = A_ - 1; } else { *lo = last_bid + 1; *hi = A_; }
  }
  // normalize_beliefs_inplace (recursive_solving.cc:41-44, util.h:68-78)
  void normalize(double* b) const {
    double s = 0;
    for (int h = 0; h < H_; ++h) s += b[h] + 1e-80;
    for (int h = 0; h < H_; ++h) b[h] = (b[h] + 1e-80) / s;
  }
  // policy[hand][action] of node `n` as the reference sees it: dense over all actions, zeros outside the legal range
  double sigmaAt(const std::vector<cfrb_node>& t, const double* sigma, int n, int hand, int action) const {
    int lo, hi; bidRange(t[n].last_bid, &lo, &hi);
    if (action < lo || action >= lo + (t[n].children_end - t[n].children_begin)) return 0.0;
    const int c = t[n].children_begin + action - lo;
    return sigma[(size_t)(c - 1) * H_ + hand];
  }

  void sampleToLeaf(Game& G, const double* sigma) {   // recursive_solving.cc:192-246
    const auto& t = tree(G.last_bid);
    std::vector<std::pair<int, int>> path;
    int node = 0;
    const int br_sampler = std::uniform_int_distribution<>(0, 1)(G.gen);
    std::vector<double> sb = G.beliefs;
    std::vector<double> policy(A_);
    while (t[node].children_end - t[node].children_begin) {
      const float eps = std::uniform_real_distribution<float>(0, 1)(G.gen);
      const int pid = G.player ^ (t[node].depth & 1);
      int lo, hi; bidRange(t[node].last_bid, &lo, &hi);
      int action;
      if (pid == br_sampler && eps < cfg_.random_action_prob) {
        action = std::uniform_int_distribution<>(lo, hi - 1)(G.gen);
      } else {
        std::discrete_distribution<> hand_dis(sb.begin() + (size_t)pid * H_, sb.begin() + (size_t)(pid + 1) * H_);
        const int hand = hand_dis(G.gen);
        for (int a = 0; a < A_; ++a) policy[a] = sigmaAt(t, sigma, node, hand, a);
        std::discrete_distribution<> action_dis(policy.begin(), policy.end());
        action = action_dis(G.gen);
      }
      for (int h = 0; h < H_; ++h) sb[(size_t)pid * H_ + h] *= sigmaAt(t, sigma, node, h, action);
      normalize(sb.data() + (size_t)pid * H_);
      path.emplace_back(node, action);
      node = t[node].children_begin + action - lo;
    }
    for (auto [n, action] : path) {   // second pass with the belief-propagation strategy (== sampling strategy for CFR)
      int lo, hi; bidRange(G.last_bid, &lo, &hi);
      for (int h = 0; h < H_; ++h) G.beliefs[(size_t)G.player * H_ + h] *= sigmaAt(t, sigma, n, h, action);
      normalize(G.beliefs.data() + (size_t)G.player * H_);
      const int child = t[n].children_begin + action - lo;
      G.last_bid = t[child].last_bid;
      G.player = G.player ^ 1;
    }
  }

  void sampleSingle(Game& G, const double* sigma) {   // recursive_solving.cc:248-275
    const auto& t = tree(G.last_bid);
    const int br_sampler = std::uniform_int_distribution<>(0, 1)(G.gen);
    const float eps = std::uniform_real_distribution<float>(0, 1)(G.gen);
    int lo, hi; bidRange(G.last_bid, &lo, &hi);
    int action;
    if (G.player == br_sampler && eps < cfg_.random_action_prob) {
      action = std::uniform_int_distribution<>(lo, hi - 1)(G.gen);
    } else {
      std::discrete_distribution<> hand_dis(G.beliefs.begin() + (size_t)G.player * H_, G.beliefs.begin() + (size_t)(G.player + 1) * H_);
      const int hand = hand_dis(G.gen);
      std::vector<double> policy(A_);
      for (int a = 0; a < A_; ++a) policy[a] = sigmaAt(t, sigma, 0, hand, a);
      std::discrete_distribution<> action_dis(policy.begin(), policy.end());
      action = action_dis(G.gen);
    }
    for (int h = 0; h < H_; ++h) G.beliefs[(size_t)G.player * H_ + h] *= sigmaAt(t, sigma, 0, h, action);
    normalize(G.beliefs.data() + (size_t)G.player * H_);
    G.last_bid = action;
    G.player ^= 1;
  }

  const liars_dice::RecursiveSolvingParams cfg_;
  const int K_;
  cfrb_handle* h_ = nullptr;
  int A_ = 0, H_ = 0, Q_ = 0, stride_ = 0;
  std::vector<std::vector<cfrb_node>> trees_;
  std::vector<Game> games_;
  std::vector<int32_t> last_bid_, player_, act_;
  std::vector<double> beliefs_, snap_;
  std::vector<float> ex_q_, ex_v_;
  int64_t subgames_solved_ = 0;
};

}  // namespace rela
