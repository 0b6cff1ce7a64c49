// Parameter structs mirrored to Python by the `rela` module, field for field like the reference's
// liars_dice::SubgameSolvingParams (subgame_solving.h:43-58) and RecursiveSolvingParams (recursive_solving.h:31-38).
// cfvpy/selfplay.py fills them with setattr from cfg.env (selfplay.py:587-610) and raises on unknown keys, so the
// B200-specific knobs below are ordinary extra fields with defaults: they can be set as `env.concurrent_games=...`
// without touching selfplay.py, and are invisible to configs that do not mention them.
#pragma once
#include <cstdlib>

namespace liars_dice {

struct SubgameSolvingParams {
  int num_iters = 10;
  int max_depth = 2;
  bool linear_update = false;
  bool use_cfr = false;   // false = fictitious play (the YAML default), true = CFR; both run on the GPU
  bool optimistic = false;
  bool dcfr = false;
  double dcfr_alpha = 0;
  double dcfr_beta = 0;
  double dcfr_gamma = 0;
};

inline int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v && *v ? std::atoi(v) : dflt;
}

struct RecursiveSolvingParams {
  int num_dice = 0;
  int num_faces = 0;
  float random_action_prob = 1.0f;
  bool sample_leaf = false;
  SubgameSolvingParams subgame_params;
  // ---- rebel_b200 extensions
  int concurrent_games = env_int("CFRB_CONCURRENT_GAMES", 1024);   // self-play games advanced in lock-step per thread loop
  int net_mode = env_int("CFRB_NET_MODE", 3);                      // include/cfrb200.h CFRB_NET_*: 3 = tcgen05 fp16 operands, fast tanh GELU
  int state_dtype = env_int("CFRB_STATE_DTYPE", 0);                // CFRB_STATE_*: 0 = fp64 tables
  int host_walk = env_int("CFRB_HOST_WALK", 0);                    // 1 = per-game sampling on the host (parity mode of the device walk)
};

// The tensor-core value net packs at most 16 output features (num_hands <= 16); larger games run the fp32 SIMT net.
inline int effective_net_mode(const RecursiveSolvingParams& p) {
  long hands = 1;
  for (int i = 0; i < p.num_dice; ++i) hands *= p.num_faces;
  return (p.net_mode >= 2 && hands > 16) ? 1 : p.net_mode;
}

}  // namespace liars_dice
