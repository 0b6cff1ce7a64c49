"""Data-generation throughput through the drop-in `rela` module exactly as cfvpy/selfplay.py drives it (initialize_datagen,
selfplay.py:182-260; benchmark loop :285-293): ModelLocker + ValuePrioritizedReplay + create_cfr_thread x threads_per_gpu +
Context.  Metric: replay.num_add()/2 x num_iters / seconds = CFR subgame-iters/s, mixed subgames of real games."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rebel_b200.rela as rela
from rebel_b200.models import make_selfplay_net


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dice", type=int, default=1)
    ap.add_argument("--faces", type=int, default=6)
    ap.add_argument("--games", type=int, default=8192, help="concurrent games per thread loop")
    ap.add_argument("--threads", type=int, default=2, help="thread loops per GPU (selfplay.threads_per_gpu)")
    ap.add_argument("--iters", type=int, default=1024)
    ap.add_argument("--fp", action="store_true", help="fictitious play instead of CFR (use_cfr = false, the YAML default)")
    ap.add_argument("--devices", type=int, default=1, help="GPUs: one ModelLocker per cuda:i like selfplay.py:193-220")
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--warmup", type=float, default=4.0)
    args = ap.parse_args()
    D, F = args.dice, args.faces
    ref_models = [[torch.jit.script(make_selfplay_net(D, F))] for _ in range(args.devices)]
    lockers = [rela.ModelLocker(m, f"cuda:{i}") for i, m in enumerate(ref_models)]
    replay = rela.ValuePrioritizedReplay(capacity=1 << 22, seed=10001, alpha=1.0, beta=1.0, prefetch=8, use_priority=False,
                                         compressed_values=False)
    cfg = rela.RecursiveSolvingParams()
    cfg.num_dice, cfg.num_faces, cfg.random_action_prob, cfg.sample_leaf = D, F, 0.25, True
    cfg.concurrent_games = args.games
    sp = cfg.subgame_params
    sp.num_iters, sp.max_depth, sp.linear_update, sp.use_cfr = args.iters, 2, True, not args.fp
    ctx = rela.Context()
    for d, locker in enumerate(lockers):
        for i in range(args.threads):
            ctx.push_env_thread(rela.create_cfr_thread(locker, replay, cfg, d * 1000 + i))
    ctx.start()
    time.sleep(args.warmup)
    n0, t0 = replay.num_add(), time.time()
    while time.time() - t0 < args.seconds:
        time.sleep(0.25)
        if replay.size() > (1 << 21):
            replay.pop_until(1 << 20)
    n1, t1 = replay.num_add(), time.time()
    ctx.terminate()
    while not ctx.terminated():
        time.sleep(0.05)
    ex = (n1 - n0) / (t1 - t0)
    print(json.dumps({"game": f"{D}x{F}f", "concurrent_games": args.games, "devices": args.devices, "thread_loops_per_device": args.threads, "solver": "fp" if args.fp else "cfr", "iters": args.iters,
                      "examples_per_s": ex, "subgames_per_s": ex / 2, "subgame_iters_per_s": ex / 2 * args.iters,
                      "seconds": t1 - t0, "error": ctx.error()}))


if __name__ == "__main__":
    main()
