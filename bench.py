#!/usr/bin/env python
"""Benchmark of the hot path: game-ticks/s of the batched Optimax Rogue updater.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--games-per-gpu G] [--impl b200|reference]

A "step" is one tick (one ``Updater.update``, optimax_rogue/logic/updater.py:76-162) for every game
of one batch of G games per GPU, with auto-reset and on-device level generation
(BASELINE.json configs[3]: 1M games, reference-exact ruleset R0, RandomBot-style uniform command
streams). Rank r owns games [r*G, (r+1)*G) (weak scaling, no collective on the step path).

Timed regions (CUDA events on the launching stream, max over ranks):
  value     K steps captured in one CUDA graph, commands already resident in HBM
  e2e       the public ``BatchedUpdater.host_stepper`` call with pinned HOST command/result buffers: the
            step's commands cross PCIe host->device and its results device->host inside the timed
            region every step (the tick kernel's TMA producer reads/writes the pinned buffers
            directly, tile by tile), then a stream sync so the caller can read the results. Headline:
            nibble-packed commands (1 B per game in, 1 B out); ``unpacked`` = uint8[N,2] commands;
            ``pipelined`` = two independent batches in flight, no per-step stream sync
  rollout   (extra) fused multi-tick kernel with both bots on device
  roofline.large_batch  (extra) the value leg at 4 G games per launch, where the fixed cost of a launch weighs a quarter
L2: the timed loop rotates over B independent batches whose combined state exceeds the 126 MB L2.

``--impl reference`` times the CPU side instead: the plain-C restatement of the reference updater
(oracle/orx_oracle.c, kind "port" -- the Python reference itself cannot travel to the GPU box) with
OpenMP over all host cores on the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'game-ticks/sec (whole box), 1M games per B200, ruleset R0'
UNIT = 'game-ticks/s'
WORKLOAD = 'configs[3]: 2^20 concurrent games per GPU, 60x10 EmptyDungeon levels generated on device, ' \
           'auto-reset, max_ticks=1000, uniform random commands (RandomBot vs RandomBot)'
MAX_TICKS = 1000
SEED = 0x0A11CE
B_ALG = 61   # bytes per game-tick: 2 x 29 B state planes + 2 B commands + 1 B result


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--games-per-gpu', type=int, default=1 << 20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--rollout-ticks', type=int, default=64)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers
def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def recorded_traffic():
    """dram bytes per launch of k_step from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(smax), 'reasons': sorted(reasons), 'samples': len(sm)}


def sim_config(auto_reset=True):
    from optimax_rogue_b200 import SimConfig
    return SimConfig(width=60, height=10, max_ticks=MAX_TICKS, seed=SEED, auto_reset=auto_reset)


# ------------------------------------------------------------------------------------------ CPU side
def cpu_port_run_isolated(n_games, seconds):
    """cpu_port_run in a child process whose environment lets OpenMP use every host core (launchers
    such as torchrun export OMP_NUM_THREADS=1, which libgomp honours even after omp_set_num_threads
    once it has been initialised that way inside a Python process)."""
    env = {k: v for k, v in os.environ.items() if not k.startswith('OMP_') and k != 'GOMP_CPU_AFFINITY'}
    code = ('import json, sys; sys.path.insert(0, %r); import bench; '
            'print(json.dumps(bench.cpu_port_run(%d, %f)))' % (ROOT, n_games, seconds))
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, check=True).stdout
    return tuple(json.loads(out.strip().splitlines()[-1]))


def cpu_port_run(n_games, seconds, steps=None, game_id_base=0):
    """Times the C restatement (OpenMP, all cores) ticking ``n_games`` games with RandomBot commands.
    Returns (ticks_per_s, cores, ticks_done, elapsed)."""
    import numpy as np
    from oracle import cport
    cfg = sim_config()
    cores = cport.set_threads(os.cpu_count())         # torchrun exports OMP_NUM_THREADS=1
    orc = cport.Oracle(cfg, n_games, game_id_base)
    orc.reset()
    stats = np.zeros(8, np.uint64)
    orc.rollout(1, 1, 1, stats)                       # warm: page in, spin up the thread pool
    stats[:] = 0
    t0 = time.perf_counter()
    done = 0
    while True:
        orc.rollout(1, 1, 1, stats)
        done += 1
        el = time.perf_counter() - t0
        if (steps is not None and done >= steps) or (steps is None and el >= seconds):
            break
    el = time.perf_counter() - t0
    return float(stats[0]) / el, cores, int(stats[0]), el


def run_reference(args, rank, world):
    """--impl reference: rank 0 alone, host cores only."""
    if rank != 0:
        return
    n = args.games_per_gpu * world      # the reference arm has no GPUs to shard over: whole-box batch on the host
    n_sample = min(n, 1 << 20)
    t_w0 = time.perf_counter()
    import numpy as np
    from oracle import cport
    cfg = sim_config()
    cores = cport.set_threads(os.cpu_count())         # torchrun exports OMP_NUM_THREADS=1
    orc = cport.Oracle(cfg, n_sample, 0)
    orc.reset()
    stats = np.zeros(8, np.uint64)
    for _ in range(max(args.warmup, 1)):
        orc.rollout(1, 1, 1, stats)
    stats[:] = 0
    steps = min(args.steps, 200)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.rollout(1, 1, 1, stats)
    el = time.perf_counter() - t0
    val = float(stats[0]) / el
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': world,
        'steps': steps, 'warmup': max(args.warmup, 1), 'ms_per_step': 1e3 * el / steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'int32',
        'data': 'synthetic', 'config': {'workload': WORKLOAD, 'games_per_step': n_sample},
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': f'{steps} steps x {n_sample} games, oracle/orx_oracle.c oro_rollout(1 tick, RandomBot x2), OpenMP'},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.perf_counter() - t_w0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU side
def run_b200(args, rank, local_rank, world):
    # Rank 0 must print exactly ONE line on stdout. Libraries (NCCL's version banner, for one) write to
    # file descriptor 1 directly, so point fd 1 at stderr for the duration and keep the real stdout aside.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from optimax_rogue_b200 import _lib
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator

    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the product path has no CPU fallback')
    _lib.lib()   # fail loudly if the extension is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    G, K, W = args.games_per_gpu, args.steps, args.warmup
    cfg = sim_config()
    # rotating batches: combined state must exceed L2 so no step finds its planes cached
    state_bytes = 29 * G
    n_batches = max(2, -(-300_000_000 // (state_bytes + 3 * G)))
    n_batches = min(n_batches, 64)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, MAX_TICKS, auto_reset=True)
    batches = []
    for b in range(n_batches):
        gs = BatchedGameState(cfg, G, dev, game_id_base=(rank * n_batches + b) * G)
        reset_games(gs)
        batches.append(gs)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    n_move_sets = 16
    moves = torch.randint(1, 6, (n_move_sets, G, 2), dtype=torch.uint8, device=dev, generator=gen)
    results = [torch.empty((G,), dtype=torch.uint8, device=dev) for _ in range(n_batches)]
    stream = torch.cuda.Stream(dev)

    def step(k):
        b = k % n_batches
        upd.update(batches[b], moves[k % n_move_sets], out=results[b])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    with torch.cuda.stream(stream):
        # de-phase the batches so they are not all at the same tick (spreads resets), untimed
        for b in range(n_batches):
            upd.rollout(batches[b], 1, 1, 37 * (b + 1))
        for k in range(max(W, 3)):
            step(k)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for k in range(K):
                step(k)
        graph.replay()      # one untimed replay (graph upload)
        torch.cuda.synchronize(dev)

        # ---- value: K steps, commands resident in HBM
        barrier()
        t_wall0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        graph.replay()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        t_wall1 = time.time()
        barrier()
        ms_total = e0.elapsed_time(e1)

        # ---- e2e: public API with host buffers, every step H2D + tick + D2H + sync
        from optimax_rogue_b200.logic.moves import pack_moves
        host_moves = [torch.empty((G, 2), dtype=torch.uint8, pin_memory=True) for _ in range(4)]
        host_cmds = [torch.empty((G,), dtype=torch.uint8, pin_memory=True) for _ in range(4)]     # nibble-packed
        for hm, hc in zip(host_moves, host_cmds):
            m = moves[0].cpu()
            hm.copy_(m)
            hc.copy_(pack_moves(m[:, 0], m[:, 1]))
        host_res = [torch.empty((G,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        k_e2e = max(10, min(K, 200))

        def time_host_loop(cmd_bufs, sync):
            # one bound stepper per (batch, command buffer): BatchedUpdater.host_stepper is the public call for
            # host-side loops; each step() = H2D commands + tick + D2H results (+ stream sync when sync=True)
            steppers = [upd.host_stepper(batches[k % n_batches], cmd_bufs[k % 4], host_res[k % 2], sync=sync)
                        for k in range(min(k_e2e, 4 * n_batches))]
            evs = [torch.cuda.Event(), torch.cuda.Event()]
            for k in range(3):
                steppers[k % len(steppers)]()
            torch.cuda.synchronize(dev)
            barrier()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(stream)
            if sync:
                for k in range(k_e2e):
                    steppers[k % len(steppers)]()  # synchronous: the caller can read host_res after each call
            else:
                # two independent batches in flight: step k is enqueued, then the host waits for step k-1's
                # results (its own result buffer) -- every step's results still reach the host
                for k in range(k_e2e):
                    steppers[k % len(steppers)]()
                    evs[k & 1].record(stream)
                    if k:
                        evs[(k - 1) & 1].synchronize()
                evs[(k_e2e - 1) & 1].synchronize()
            eb.record(stream)
            torch.cuda.synchronize(dev)
            barrier()
            return ea.elapsed_time(eb)

        ms_e2e_unpacked = time_host_loop(host_moves, True)
        ms_e2e = time_host_loop(host_cmds, True)
        ms_e2e_pipelined = time_host_loop(host_cmds, False)

        # ---- step + observe (extra): the self-play tick, one pass (orx_step_observe); observation
        # buffers rotate with the batches so that their writes cannot be absorbed by the L2 either
        obs_bufs = [torch.empty((G, 2, 12), dtype=torch.int16, device=dev) for _ in range(n_batches)]
        k_so = max(n_batches, min(K, 10 * n_batches))
        for k in range(3):
            upd.update_observe(batches[k % n_batches], moves[k % len(moves)], stairs_radius=4, out=results[k % n_batches],
                               obs_out=obs_bufs[k % n_batches])
        torch.cuda.synchronize(dev)
        g_so = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_so, stream=stream):
            for k in range(k_so):
                upd.update_observe(batches[k % n_batches], moves[k % len(moves)], stairs_radius=4,
                                   out=results[k % n_batches], obs_out=obs_bufs[k % n_batches])
        g_so.replay()
        torch.cuda.synchronize(dev)
        barrier()
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es0.record(stream)
        g_so.replay()
        es1.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        ms_so = es0.elapsed_time(es1)
        del obs_bufs, g_so

        # ---- rollout (extra): fused T-tick kernel, both bots on device
        T = args.rollout_ticks
        stats = torch.zeros((8,), dtype=torch.int64, device=dev)
        r_launches = max(2, min(n_batches, 8))
        for b in range(2):
            upd.rollout(batches[b], 1, 1, T, stats)
        torch.cuda.synchronize(dev)
        stats.zero_()
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record(stream)
        for b in range(r_launches):
            upd.rollout(batches[b % n_batches], 1, 1, T, stats)
        e5.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        ms_roll = e4.elapsed_time(e5)
        roll_ticks = int(stats[0].item())

        # ---- the same tick kernel at 4x the batch (extra): the fixed per-launch cost (ramp / drain of one
        # dependent kernel boundary, ~2.4 us) weighs a quarter as much, which shows what the kernel body streams
        GL, nbl, KL = 4 * G, 3, 60
        big = []
        for b in range(nbl):
            gsl = BatchedGameState(cfg, GL, dev, game_id_base=(1 << 40) + (rank * nbl + b) * GL)
            reset_games(gsl)
            upd.rollout(gsl, 1, 1, 29 * (b + 1))
            big.append(gsl)
        big_moves = torch.randint(1, 6, (4, GL, 2), dtype=torch.uint8, device=dev, generator=gen)
        big_res = torch.empty((GL,), dtype=torch.uint8, device=dev)
        for k in range(3):
            upd.update(big[k % nbl], big_moves[k % 4], out=big_res)
        torch.cuda.synchronize(dev)
        g_big = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_big, stream=stream):
            for k in range(KL):
                upd.update(big[k % nbl], big_moves[k % 4], out=big_res)
        g_big.replay()
        torch.cuda.synchronize(dev)
        barrier()
        e8, e9 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e8.record(stream)
        g_big.replay()
        e9.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        ms_big = e8.elapsed_time(e9)
        del big, big_moves, big_res, g_big

        # ---- strong scaling (extra, N > 1): BASELINE.json's "1M games sharded across 8xB200" read literally,
        # 2^20 games in total, 2^20 / N per GPU; still one launch per step per GPU, rotating batches > L2
        ms_strong, Gs, KS = 0.0, G // world, 400
        if world > 1:
            nbs = min(64, max(2, -(-300_000_000 // (32 * Gs))))
            small = []
            for b in range(nbs):
                gss = BatchedGameState(cfg, Gs, dev, game_id_base=(1 << 41) + b * G + rank * Gs)
                reset_games(gss)
                upd.rollout(gss, 1, 1, 11 * (b + 1))
                small.append(gss)
            small_res = torch.empty((Gs,), dtype=torch.uint8, device=dev)
            for k in range(3):
                upd.update(small[k % nbs], moves[k % n_move_sets][:Gs], out=small_res)
            torch.cuda.synchronize(dev)
            g_small = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_small, stream=stream):
                for k in range(KS):
                    upd.update(small[k % nbs], moves[k % n_move_sets][:Gs], out=small_res)
            g_small.replay()
            torch.cuda.synchronize(dev)
            barrier()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(stream)
            g_small.replay()
            eb.record(stream)
            torch.cuda.synchronize(dev)
            barrier()
            ms_strong = ea.elapsed_time(eb)
            del small, small_res, g_small

        # ---- ruleset R1 (extra; README-only rules, parity unpinned): configs[2], 65,536 games per GPU
        from optimax_rogue_b200.r1 import R1GameState
        G1, nb1, K1 = 1 << 16, 18, 72
        r1_batches = [R1GameState(G1, max_ticks=MAX_TICKS, auto_reset=True, seed=SEED, device=dev,
                                  game_id_base=(rank * nb1 + b) * G1).reset() for b in range(nb1)]
        r1_moves = torch.randint(1, 7, (4, G1, 2), dtype=torch.uint8, device=dev, generator=gen)
        r1_res = torch.empty((G1,), dtype=torch.uint8, device=dev)
        for b in range(nb1):
            r1_batches[b].rollout(64)            # populate with enemies / items, untimed
        torch.cuda.synchronize(dev)
        g1 = torch.cuda.CUDAGraph()
        for k in range(3):
            r1_batches[k].update(r1_moves[k % 4], out=r1_res)
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(g1, stream=stream):
            for k in range(K1):
                r1_batches[k % nb1].update(r1_moves[k % 4], out=r1_res)
        g1.replay()
        torch.cuda.synchronize(dev)
        barrier()
        e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e6.record(stream)
        g1.replay()
        e7.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        ms_r1 = e6.elapsed_time(e7)
        r1_stats = torch.zeros((8,), dtype=torch.int64, device=dev)
        e8, e9 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e8.record(stream)
        for b in range(4):
            r1_batches[b].rollout(64, r1_stats)
        e9.record(stream)
        torch.cuda.synchronize(dev)
        barrier()
        ms_r1_roll = e8.elapsed_time(e9)

    if rank == 0:
        time.sleep(0.15)
        sampler.stop()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_total = max_over_ranks(ms_total)
    ms_e2e = max_over_ranks(ms_e2e)
    ms_e2e_unpacked = max_over_ranks(ms_e2e_unpacked)
    ms_e2e_pipelined = max_over_ranks(ms_e2e_pipelined)
    ms_roll = max_over_ranks(ms_roll)
    ms_so = max_over_ranks(ms_so)
    ms_big = max_over_ranks(ms_big)
    ms_strong = max_over_ranks(ms_strong)
    ms_r1 = max_over_ranks(ms_r1)
    ms_r1_roll = max_over_ranks(ms_r1_roll)
    if world > 1:
        t = torch.tensor([roll_ticks], dtype=torch.int64, device=dev)
        dist.all_reduce(t)          # the optional end-of-rollout stats gather (tiny, off the step path)
        roll_ticks_all = int(t.item())
    else:
        roll_ticks_all = roll_ticks

    if rank == 0:
        peak, peak_src = measured_peak()
        ms_step = ms_total / K
        value = world * G * K / (ms_total * 1e-3)
        achieved = B_ALG * G / (ms_step * 1e-3) / 1e9
        tr = recorded_traffic()
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'games_per_gpu': G, 'global_games': G * world,
                       'parallelism': f'{world} shard(s) of independent games, no collective on the step path',
                       'l2': f'rotating {n_batches} independent batches per GPU ({n_batches * (state_bytes + 3 * G) / 1e6:.0f} MB of planes > 126 MB L2), no flush needed',
                       'launch': 'K steps captured in one CUDA graph'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': (tr or {}).get('dram_bytes_per_launch'), 'peak_source': peak_src,
                         'kernel': 'k_step_pipe<EMPTY,CMD_BYTES>', 'alg_bytes_per_game_tick': B_ALG,
                         'games_per_launch': G,
                         'large_batch': {'games_per_launch': GL, 'us_per_step': ms_big / KL * 1e3, 'steps': KL,
                                         'achieved': B_ALG * GL / (ms_big / KL * 1e-3) / 1e9,
                                         'frac': B_ALG * GL / (ms_big / KL * 1e-3) / 1e9 / peak,
                                         'note': 'same kernel, same config, 4x the games per launch (3 rotating batches, '
                                                 f'{3 * 32 * GL / 1e6:.0f} MB of planes): the per-launch ramp/drain weighs a quarter as much'}},
            'e2e': {'value': world * G * k_e2e / (ms_e2e * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': G, 'd2h_bytes_per_step': G, 'steps': k_e2e,
                    'api': 'BatchedUpdater.host_stepper(state, pinned host uint8[N] commands p1|p2<<4, pinned host uint8[N] results)() '
                           '= orx_step_host_packed_sync: commands and results cross PCIe inside the tick kernel, stream sync every step',
                    'unpacked': {'value': world * G * k_e2e / (ms_e2e_unpacked * 1e-3), 'h2d_bytes_per_step': 2 * G,
                                 'd2h_bytes_per_step': G, 'api': 'same call with uint8[N,2] commands (orx_step_host_sync)'},
                    'pipelined': {'value': world * G * k_e2e / (ms_e2e_pipelined * 1e-3), 'h2d_bytes_per_step': G,
                                  'd2h_bytes_per_step': G,
                                  'api': 'host_stepper(..., sync=False) = orx_step_host_packed on two independent batches in flight; '
                                         'the host waits on step k-1\'s event after enqueueing step k'}},
            'gpu_launches': K,
            **({'strong_scaling': {'value': world * Gs * KS / (ms_strong * 1e-3), 'unit': UNIT, 'global_games': world * Gs,
                                   'games_per_gpu': Gs, 'us_per_step': ms_strong / KS * 1e3, 'steps': KS,
                                   'note': '2^20 games in total sharded over the GPUs (BASELINE.json configs[3] read literally); '
                                           'the headline value is the weak-scaling figure at 2^20 games per GPU'}} if world > 1 else {}),
            'step_observe': {'value': world * G * k_so / (ms_so * 1e-3), 'unit': UNIT, 'us_per_step': ms_so / k_so * 1e3,
                             'steps': k_so, 'alg_bytes_per_game_tick': B_ALG + 48,
                             'hbm_frac': (B_ALG + 48) * G / (ms_so / k_so * 1e-3) / 1e9 / peak,
                             'note': 'orx_step_observe: the tick plus both players\' observations (int16[N,2,12]) '
                                     'of the resulting state in one pass; observation buffers rotate with the batches'},
            'rollout': {'value': roll_ticks_all / (ms_roll * 1e-3), 'unit': UNIT, 'ticks_per_launch': T,
                        'launches': r_launches, 'fused': True,
                        'note': 'orx_rollout: bots + tick fused, state in registers for T ticks'},
            'r1': {'note': 'ruleset R1 = README-only rules (docs/RULESET_R1.md); PARITY UNPINNED vs the reference, '
                           'bit-exact vs oracle/orx_r1_oracle.c; configs[2]: 65,536 games per GPU, 8 enemy + 4 item slots',
                   'value': world * G1 * K1 / (ms_r1 * 1e-3), 'unit': UNIT, 'games_per_gpu': G1, 'steps': K1,
                   'us_per_step': ms_r1 / K1 * 1e3, 'alg_bytes_per_game_tick': 2 * 241 + 3,
                   'hbm_frac': (2 * 241 + 3) * G1 / (ms_r1 / K1 * 1e-3) / 1e9 / peak,
                   'rollout_value': world * 4 * G1 * 64 / (ms_r1_roll * 1e-3)},
            'clocks': sampler.summary(t_wall0, t_wall1),
        }
        if not args.no_cpu_baseline and world == 1:     # CPU baseline: rank 0 at N=1 only (other ranks would spin on host cores)
            v, cores, ticks, el = cpu_port_run_isolated(1 << 18, args.cpu_seconds)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': f'{ticks} game-ticks in {el:.1f} s: 2^18 games, same config, oracle/orx_oracle.c oro_rollout (RandomBot x2), OpenMP over all host cores'}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1 and args.gpus > 1 and 'RANK' not in os.environ:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29517', os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == 'reference':
        if rank == 0 and any(k.startswith('OMP_') for k in os.environ) and '_ORX_BENCH_CHILD' not in os.environ:
            # torchrun exports OMP_NUM_THREADS=1: re-run rank 0's measurement with every host core
            env = {k: v for k, v in os.environ.items() if not k.startswith('OMP_')}
            env['_ORX_BENCH_CHILD'] = '1'
            sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
