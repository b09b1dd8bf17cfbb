#!/usr/bin/env python
"""Benchmark of the hot path: game-ticks/s of the batched Optimax Rogue updater.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--games G] [--impl b200|reference]

A "step" is one tick (one ``Updater.update``, optimax_rogue/logic/updater.py:76-162) for every game of the
batch: BASELINE.json configs[3], G = 2^20 concurrent games of the reference-exact ruleset R0 with on-device
level generation and auto-reset, uniform random commands (RandomBot vs RandomBot), sharded over the N GPUs of
the box by global game id (rank r owns games [r*G/N, (r+1)*G/N), no collective on the step path). The total
is fixed, so the line says ``"scaling": "strong"``; at N > 1 the weak-scaling figure (2^20 games per GPU) is
reported beside it as ``weak_scaling``.

Timed regions (CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks).
Every region lasts at least ~50 ms: a K-step CUDA graph is replayed back to back inside ONE event pair as
often as that takes (``replays``), and the mean step time is reported.
  value     K steps captured in one CUDA graph, commands already resident in HBM
  e2e       the public ``BatchedUpdater.host_stepper`` call with pinned HOST command/result buffers: each
            step's commands cross PCIe host->device and its results device->host inside the timed region,
            then a stream sync so the caller can read the results -- every step
  extras    roofline.large_batch (4x the games per launch), roofline.grid_wait_mode (the default ordering), step_observe, rollout, config2 (BASELINE.json
            configs[1]: 4,096 games on one fixed wall map), r1 (README-only ruleset, parity unpinned)
L2: each leg rotates over B independent batches whose combined state exceeds the 126 MB L2.

CPU side. ``cpu_baseline`` (rank 0, N = 1 only, in a child process after the GPU legs) and ``--impl reference``
time the reference's OWN Python updater loop (optimax_rogue/server/main.py:110-113) on all host cores when the
reference tree is present (``/root/reference`` in the build container, ``oracle/_ref`` -- shipped by
``oracle/make_ref.py`` -- on the GPU box): ``kind: "reference"``. The plain-C restatement
(oracle/orx_oracle.c, OpenMP) is timed as well (``cpu_baseline.port``) and is the fallback, labelled
``kind: "port"``, when no reference tree exists.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'game-ticks/sec (whole box) at 1/2/4/8 B200, 1M games; % HBM roofline'
UNIT = 'game-ticks/s'
WORKLOAD = 'configs[3]: 2^20 concurrent games sharded over the GPUs, ruleset R0, 60x10 EmptyDungeon levels generated ' \
           'on device, auto-reset, max_ticks=1000, uniform random commands (RandomBot vs RandomBot)'
MAX_TICKS = 1000
SEED = 0x0A11CE
B_ALG = 61   # bytes per game-tick: 2 x 29 B state planes + 2 B commands + 1 B result
MIN_WINDOW_MS = 50.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--games', type=int, default=1 << 20, help='games in the whole job (sharded over the GPUs)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--rollout-ticks', type=int, default=64)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='only value + e2e (+ weak_scaling)')
    ap.add_argument('--cpu-seconds', type=float, default=8.0)
    return ap.parse_args()


def workload_config(args, world):
    """The ``config`` object: identical, key for key and value for value, in both arms."""
    return {'workload': WORKLOAD, 'global_games': args.games, 'games_per_gpu': args.games // world,
            'parallelism': f'{world} shard(s) of independent games by global game id, no collective on the step path',
            'max_ticks': MAX_TICKS, 'width': 60, 'height': 10}


# ------------------------------------------------------------------------------------------ helpers
def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def recorded_traffic():
    """dram bytes per launch of the tick kernel from the committed ncu --set full capture, or None."""
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


def sim_config(auto_reset=True, overlap_ticks=False):
    from optimax_rogue_b200 import SimConfig
    return SimConfig(width=60, height=10, max_ticks=MAX_TICKS, seed=SEED, auto_reset=auto_reset, overlap_ticks=overlap_ticks)


# ------------------------------------------------------------------------------------------ CPU side
def _clean_env():
    """Launchers such as torchrun export OMP_NUM_THREADS=1, which libgomp honours for good once a process has
    started with it: CPU measurements run in a child process without it."""
    return {k: v for k, v in os.environ.items() if not k.startswith('OMP_') and k != 'GOMP_CPU_AFFINITY'}


def _child_json(code):
    out = subprocess.run([sys.executable, '-c', code], env=_clean_env(), capture_output=True, text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def reference_available():
    from oracle import ref_harness
    return ref_harness.reference_available()


def cpu_reference_run_isolated(games, seconds=None, steps=None, warmup=1):
    """The live Python reference on all host cores (oracle/ref_pool.py), in a child process (it forks workers,
    which must not happen in a process that holds a CUDA context)."""
    code = ('import json, sys; sys.path.insert(0, %r); from oracle import ref_pool; '
            'print(json.dumps(ref_pool.timed_run(%d, steps=%r, seconds=%r, warmup=%d, max_ticks=%d)))'
            % (ROOT, games, steps, seconds, warmup, MAX_TICKS))
    return _child_json(code)


def cpu_port_run_isolated(n_games, seconds=None, steps=None, warmup=1):
    code = ('import json, sys; sys.path.insert(0, %r); import bench; '
            'print(json.dumps(bench.cpu_port_run(%d, %r, %r, %d)))' % (ROOT, n_games, seconds, steps, warmup))
    return _child_json(code)


def cpu_port_run(n_games, seconds=None, steps=None, warmup=1):
    """Times the C restatement (OpenMP, all cores) ticking ``n_games`` games with RandomBot commands."""
    import numpy as np
    from oracle import cport
    cfg = sim_config()
    cores = cport.set_threads(os.cpu_count())
    orc = cport.Oracle(cfg, n_games, 0)
    orc.reset()
    stats = np.zeros(8, np.uint64)
    for _ in range(max(warmup, 1)):
        orc.rollout(1, 1, 1, stats)                   # warm: page in, spin up the thread pool
    stats[:] = 0
    t0 = time.perf_counter()
    done = 0
    while True:
        orc.rollout(1, 1, 1, stats)
        done += 1
        el = time.perf_counter() - t0
        if (steps is not None and done >= steps) or (steps is None and el >= seconds):
            break
    return {'value': float(stats[0]) / el, 'cores': cores, 'games': n_games, 'steps': done, 'ticks': int(stats[0]), 'elapsed': el}


def reference_sample_games(total):
    """Games per step of the CPU arms: a bounded sample of the workload, about 2,048 games per host core, so
    that a step of the Python reference takes ~60 ms whatever the core count."""
    return max(256, min(total, 2048 * (os.cpu_count() or 1)))


def run_reference(args, rank, world):
    """--impl reference: rank 0 alone, host cores only."""
    if rank != 0:
        return
    t_w0 = time.perf_counter()
    W = max(args.warmup, 1)
    sample = reference_sample_games(args.games)
    if reference_available():
        r = cpu_reference_run_isolated(sample, steps=args.steps, warmup=W)
        kind = 'reference'
        what = (f'{r["steps"]} steps x {r["games"]} games: the unmodified reference loop (GameState.on_tick + Updater.update + 2 x RandomBot.move, '
                f'optimax_rogue/server/main.py:110-113), natively seeded, one process per host core, Python {r["python"]} numpy {r["numpy"]}')
    else:
        r = cpu_port_run_isolated(min(args.games, 1 << 20), steps=args.steps, warmup=W)
        kind = 'port'
        what = f'{r["steps"]} steps x {r["games"]} games: oracle/orx_oracle.c oro_rollout(1 tick, RandomBot x2), OpenMP; no reference tree on this machine'
    val = r['value']
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': world,
        'steps': r['steps'], 'warmup': W, 'ms_per_step': 1e3 * r['elapsed'] / r['steps'],
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'int32',
        'data': 'synthetic', 'config': workload_config(args, world),
        'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': r['cores'], 'kind': kind, 'sample': what},
        'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'sample_games_per_step': r['games'], 'gpu_launches': 0, 'wall_s': time.perf_counter() - t_w0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU side
def run_b200(args, rank, local_rank, world):
    # Rank 0 must print exactly ONE line on stdout. Libraries (NCCL's version banner, for one) write to
    # file descriptor 1 directly, so point fd 1 at stderr for the duration and keep the real stdout aside.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    from optimax_rogue_b200 import SimConfig, _abi, _lib
    from optimax_rogue_b200.game.state import BatchedGameState
    from optimax_rogue_b200.logic.moves import pack_moves
    from optimax_rogue_b200.logic.updater import BatchedUpdater, reset_games
    from optimax_rogue_b200.logic.worldgen import EmptyDungeonGenerator, FixedDungeonGenerator

    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the product path has no CPU fallback')
    _lib.lib()   # fail loudly if the extension is missing
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa = None
    if world > 1:
        from optimax_rogue_b200.parallel import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local_rank)      # before any pinned host buffer exists: first touch on the GPU's node
        dist.init_process_group('nccl', device_id=dev)

    K, W = args.steps, max(args.warmup, 3)
    G_total = args.games
    G = G_total // world                     # this rank's shard (strong scaling: the total is fixed)
    # Every device-timed leg enqueues its ticks back to back (K steps in one CUDA graph): the states opt into the
    # throughput mode (SimConfig.overlap_ticks = ORX_PATH_TILE_FLAGS), in which consecutive tick launches overlap run
    # by run (a run = the tiles of one CTA). The e2e legs synchronise after every tick; their *_sync entry points run in grid-wait mode regardless.
    cfg = sim_config(overlap_ticks=True)
    upd = BatchedUpdater(EmptyDungeonGenerator(60, 10), 1, MAX_TICKS, auto_reset=True)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    stream = torch.cuda.Stream(dev)
    launches = {'n': 0}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def n_batches_for(games, bytes_per_game=32):
        # rotating batches: their combined state must exceed the L2 so that no step finds its planes cached
        return min(64, max(2, -(-300_000_000 // (bytes_per_game * games))))

    def make_batches(games, nb, gid0, c=cfg, dephase=37):
        out = []
        for b in range(nb):
            gs = BatchedGameState(c, games, dev, game_id_base=gid0 + b * games)
            reset_games(gs)
            out.append(gs)
        for b, gs in enumerate(out):       # de-phase the batches so they are not all at the same tick (spreads resets), untimed
            if dephase:
                upd.rollout(gs, 1, 1, dephase * (b + 1) % 997 + 1)
        return out

    def timed_graph(step_fn, k_steps, nb, min_ms=MIN_WINDOW_MS):
        """Captures ``m * k_steps`` steps -- m >= 1 the smallest count that makes the graph a whole number of turns
        through the nb rotating batches, so that step k ticks batch k % nb across replays as well and no batch is
        ticked again before the nb - 1 others -- and replays the graph R times back to back inside one event pair, R
        such that the window lasts >= min_ms.
        Returns (ms per step, replays of the K-step sequence, window ms, wall-clock bounds of the window)."""
        m = nb // math.gcd(k_steps, nb)
        for k in range(W):
            step_fn(k)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for k in range(m * k_steps):
                step_fn(k)
        g.replay()      # untimed: graph upload
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); g.replay(); e1.record(stream)      # probe: how long is one replay
        torch.cuda.synchronize(dev)
        probe = max_over_ranks(e0.elapsed_time(e1))
        R = max(1, int(math.ceil(min_ms / max(probe, 1e-3))))
        barrier()
        t0 = time.time()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(R):
            g.replay()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        t1 = time.time()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches['n'] = m * k_steps * R
        del g
        return ms / (m * k_steps * R), m * R, ms, (t0, t1)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    out = {}
    with torch.cuda.stream(stream):
        # ---------------------------------------------------------------- value: the tick, commands resident in HBM
        nb = n_batches_for(G)
        batches = make_batches(G, nb, rank * nb * G)
        n_move_sets = 16
        moves = torch.randint(1, 6, (n_move_sets, G, 2), dtype=torch.uint8, device=dev, generator=gen)
        results = [torch.empty((G,), dtype=torch.uint8, device=dev) for _ in range(nb)]

        def step(k):
            b = k % nb
            upd.update(batches[b], moves[k % n_move_sets], out=results[b])

        ms_step, replays, window_ms, (t_wall0, t_wall1) = timed_graph(step, K, nb)
        value_launches = launches['n']

        # ---------------------------------------------------------------- e2e: public API, host buffers, sync every step
        n_host = 4
        host_moves = [torch.empty((G, 2), dtype=torch.uint8, pin_memory=True) for _ in range(n_host)]
        host_cmds = [torch.empty((G,), dtype=torch.uint8, pin_memory=True) for _ in range(n_host)]     # nibble-packed
        for k, (hm, hc) in enumerate(zip(host_moves, host_cmds)):
            m = moves[k].cpu()
            hm.copy_(m)
            hc.copy_(pack_moves(m[:, 0], m[:, 1]))
        host_res = [torch.empty((G,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]

        def time_host_loop(cmd_bufs, sync, res_bufs=None, bits=False, states=None):
            # one bound stepper per (batch, command buffer): BatchedUpdater.host_stepper is the public call for
            # host-side loops; each step() = H2D commands + tick + D2H results (+ stream sync when sync=True)
            res_bufs = res_bufs or host_res
            states = states or batches
            steppers = [upd.host_stepper(states[k % len(states)], cmd_bufs[k % n_host], res_bufs[k % 2], sync=sync, bits=bits)
                        for k in range(len(states) * n_host // math.gcd(len(states), n_host))]
            evs = [torch.cuda.Event(), torch.cuda.Event()]

            def run(n):
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record(stream)
                if sync:
                    for k in range(n):
                        steppers[k % len(steppers)]()  # synchronous: the caller can read host_res after each call
                else:
                    # two independent batches in flight: step k is enqueued, then the host waits for step k-1's
                    # results (its own result buffer) -- every step's results still reach the host
                    for k in range(n):
                        steppers[k % len(steppers)]()
                        evs[k & 1].record(stream)
                        if k:
                            evs[(k - 1) & 1].synchronize()
                    evs[(n - 1) & 1].synchronize()
                eb.record(stream)
                torch.cuda.synchronize(dev)
                return ea.elapsed_time(eb)

            run(5)
            probe = max_over_ranks(run(20)) / 20
            n = int(min(5000, max(50, math.ceil(MIN_WINDOW_MS / max(probe, 1e-4)))))
            barrier()
            ms = run(n)
            barrier()
            return max_over_ranks(ms) / n, n

        from optimax_rogue_b200.logic.moves import pack_moves5
        nb_in, nb_out = _abi.cmd5_bytes(G), _abi.res2_bytes(G)
        host_cmd5 = [torch.from_numpy(pack_moves5(moves[k][:, 0].cpu().numpy(), moves[k][:, 1].cpu().numpy())).pin_memory() for k in range(n_host)]
        host_res2 = [torch.empty((nb_out,), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        ms_e2e, k_e2e = time_host_loop(host_cmd5, True, res_bufs=host_res2, bits=True)
        out['e2e'] = {'value': world * G / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': nb_in, 'd2h_bytes_per_step': nb_out,
                      'steps': k_e2e, 'us_per_step': ms_e2e * 1e3,
                      'api': 'BatchedUpdater.host_stepper(state, pinned host cmd5 bytes, pinned host res2 bytes, bits=True)() = orx_step_host_bits_sync: '
                             '5 bits of command pair per game host->device and 2 bits of result per game device->host cross PCIe inside the '
                             'timed region (read / written by the tick kernel, one bulk copy per CTA each way), stream sync every step'}
        if not args.no_extras:
            ms_n, k_n = time_host_loop(host_cmds, True)
            out['e2e']['nibbles'] = {'value': world * G / (ms_n * 1e-3), 'h2d_bytes_per_step': G, 'd2h_bytes_per_step': G, 'steps': k_n,
                                     'api': 'round-1 format: uint8[N] commands p1|p2<<4 in, uint8[N] results out (orx_step_host_packed_sync)'}
            ms_u, k_u = time_host_loop(host_moves, True)
            ms_p, k_p = time_host_loop(host_cmd5, False, res_bufs=host_res2, bits=True)
            out['e2e']['unpacked'] = {'value': world * G / (ms_u * 1e-3), 'h2d_bytes_per_step': 2 * G, 'd2h_bytes_per_step': G,
                                      'steps': k_u, 'api': 'same call with uint8[N,2] commands (orx_step_host_sync)'}
            out['e2e']['pipelined'] = {'value': world * G / (ms_p * 1e-3), 'h2d_bytes_per_step': nb_in, 'd2h_bytes_per_step': nb_out, 'steps': k_p,
                                       'api': 'host_stepper(..., sync=False, bits=True) = orx_step_host_bits on two independent batches in flight; '
                                              'the host waits on step k-1\'s event after enqueueing step k'}
        del host_moves, host_cmds, host_res, host_cmd5, host_res2

        # ---------------------------------------------------------------- weak scaling (N > 1): 2^20 games per GPU
        if world > 1:
            Gw = G_total
            nbw = n_batches_for(Gw)
            wb = make_batches(Gw, nbw, (1 << 41) + rank * nbw * Gw)
            wmoves = torch.randint(1, 6, (8, Gw, 2), dtype=torch.uint8, device=dev, generator=gen)
            wres = [torch.empty((Gw,), dtype=torch.uint8, device=dev) for _ in range(nbw)]
            ms_w, rep_w, _, _ = timed_graph(lambda k: upd.update(wb[k % nbw], wmoves[k % 8], out=wres[k % nbw]), K, nbw)
            out['weak_scaling'] = {'value': world * Gw / (ms_w * 1e-3), 'unit': UNIT, 'games_per_gpu': Gw, 'global_games': world * Gw,
                                   'us_per_step': ms_w * 1e3, 'steps': K, 'replays': rep_w,
                                   'note': 'the same leg with 2^20 games PER GPU (the batch grows with the box)'}
            w_cmd5 = [torch.from_numpy(pack_moves5(wmoves[k][:, 0].cpu().numpy(), wmoves[k][:, 1].cpu().numpy())).pin_memory() for k in range(n_host)]
            w_res2 = [torch.empty((_abi.res2_bytes(Gw),), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
            ms_we, k_we = time_host_loop(w_cmd5, True, res_bufs=w_res2, bits=True, states=wb)
            out['weak_scaling']['e2e'] = {'value': world * Gw / (ms_we * 1e-3), 'unit': UNIT, 'us_per_step': ms_we * 1e3, 'steps': k_we,
                                          'h2d_bytes_per_step': _abi.cmd5_bytes(Gw), 'd2h_bytes_per_step': _abi.res2_bytes(Gw),
                                          'note': 'the e2e call (bit-packed host buffers, sync every step) with 2^20 games per GPU'}
            del wb, wmoves, wres, w_cmd5, w_res2

        if not args.no_extras:
            # ------------------------------------------------------------ step + observe: the self-play tick, one pass
            obs_bufs = [torch.empty((G, 2, 12), dtype=torch.int16, device=dev) for _ in range(nb)]
            ms_so, rep_so, _, _ = timed_graph(
                lambda k: upd.update_observe(batches[k % nb], moves[k % n_move_sets], stairs_radius=4, out=results[k % nb], obs_out=obs_bufs[k % nb]),
                min(K, 10 * nb), nb)
            out['step_observe'] = {'value': world * G / (ms_so * 1e-3), 'unit': UNIT, 'us_per_step': ms_so * 1e3,
                                   'alg_bytes_per_game_tick': B_ALG + 48, 'hbm_frac_of': (B_ALG + 48) * G / (ms_so * 1e-3) / 1e9,
                                   'note': 'orx_step_observe: the tick plus both players\' observations (int16[N,2,12]) of the resulting '
                                           'state in one pass; observation buffers rotate with the batches'}
            del obs_bufs

            # ------------------------------------------------------------ rollout: fused T-tick kernel, both bots on device
            T = args.rollout_ticks
            stats = torch.zeros((8,), dtype=torch.int64, device=dev)
            r_launches = max(2, min(nb, 8))
            for b in range(2):
                upd.rollout(batches[b], 1, 1, T, stats)
            torch.cuda.synchronize(dev)
            stats.zero_()
            barrier()
            e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e4.record(stream)
            for b in range(r_launches):
                upd.rollout(batches[b % nb], 1, 1, T, stats)
            e5.record(stream)
            torch.cuda.synchronize(dev)
            barrier()
            ms_roll = max_over_ranks(e4.elapsed_time(e5))
            roll_ticks = stats[0:1].clone()
            if world > 1:
                dist.all_reduce(roll_ticks)          # the optional end-of-rollout stats gather (tiny, off the step path)
            out['rollout'] = {'value': int(roll_ticks.item()) / (ms_roll * 1e-3), 'unit': UNIT, 'ticks_per_launch': T, 'launches': r_launches,
                              'fused': True, 'note': 'orx_rollout: bots + tick fused, state in registers for T ticks'}

            # ------------------------------------------------------------ the same tick kernel at 4x the games per launch
            GL = 4 * G
            nbl = n_batches_for(GL)
            big = make_batches(GL, nbl, (1 << 40) + rank * nbl * GL, dephase=29)
            big_moves = torch.randint(1, 6, (4, GL, 2), dtype=torch.uint8, device=dev, generator=gen)
            big_res = [torch.empty((GL,), dtype=torch.uint8, device=dev) for _ in range(nbl)]
            ms_big, rep_big, _, _ = timed_graph(lambda k: upd.update(big[k % nbl], big_moves[k % 4], out=big_res[k % nbl]), min(K, 60), nbl)
            out['large_batch'] = {'games_per_launch': GL, 'us_per_step': ms_big * 1e3, 'achieved': B_ALG * GL / (ms_big * 1e-3) / 1e9,
                                  'note': f'same kernel, same config, 4x the games per launch ({nbl} rotating batches): a launch\'s fixed cost weighs a quarter as much'}
            del big, big_moves, big_res

            # ------------------------------------------------------------ the same leg in grid-wait mode (the library's default ordering)
            gw = make_batches(G, nb, (1 << 41) + rank * nb * G, c=sim_config(overlap_ticks=False))
            ms_gw, _, _, _ = timed_graph(lambda k: upd.update(gw[k % nb], moves[k % n_move_sets], out=results[k % nb]), K, nb)
            out['grid_wait_mode'] = {'games_per_launch': G, 'us_per_step': ms_gw * 1e3, 'achieved': B_ALG * G / (ms_gw * 1e-3) / 1e9,
                                     'note': 'same kernel, same batches and graph, states WITHOUT overlap_ticks: every tick launch waits for the whole previous grid '
                                             '(what a loop with other kernels between two ticks gets; throughput mode is opt-in)'}
            del gw

            # ------------------------------------------------------------ BASELINE.json configs[1]: 4,096 games, one fixed wall map
            rng = np.random.default_rng(0)
            t = np.full((60, 10), 1, np.uint8); t[[0, -1], :] = 2; t[:, [0, -1]] = 2
            t[1:-1, 1:-1][rng.random((58, 8)) < 0.10] = 2
            c2 = SimConfig(dgen_kind=_abi.DGEN_FIXED, fixed_tiles=t, max_ticks=MAX_TICKS, seed=SEED, auto_reset=True, overlap_ticks=True)
            u2 = BatchedUpdater(FixedDungeonGenerator(t), 1, MAX_TICKS, auto_reset=True)
            G2, nb2 = 4096, 64
            b2 = []
            for b in range(nb2):
                gs = BatchedGameState(c2, G2, dev, game_id_base=(1 << 42) + (rank * nb2 + b) * G2)
                reset_games(gs)
                u2.rollout(gs, 1, 1, 13 * (b + 1))
                b2.append(gs)
            res2 = [torch.empty((G2,), dtype=torch.uint8, device=dev) for _ in range(nb2)]
            ms_c2, rep_c2, _, _ = timed_graph(lambda k: u2.update(b2[k % nb2], moves[k % n_move_sets][:G2], out=res2[k % nb2]), 1024, nb2)
            out['config2'] = {'workload': 'BASELINE.json configs[1]: 4,096 concurrent games, one shared 60x10 wall map (10% interior walls) staged in shared '
                                          'memory, players only (move / attack / stay), per GPU', 'value': world * G2 / (ms_c2 * 1e-3), 'unit': UNIT,
                              'us_per_step': ms_c2 * 1e3, 'steps': 1024, 'replays': rep_c2,
                              'l2': f'{nb2} rotating batches = {nb2 * 32 * G2 / 1e6:.0f} MB of planes: L2 resident by nature, a launch is 16 tiles'}
            del b2, res2

            # ------------------------------------------------------------ ruleset R1 (README-only rules, parity unpinned)
            from optimax_rogue_b200.r1 import R1GameState
            G1, nb1 = 1 << 16, 18
            r1_batches = [R1GameState(G1, max_ticks=MAX_TICKS, auto_reset=True, seed=SEED, device=dev, overlap_ticks=True,
                                      game_id_base=(rank * nb1 + b) * G1).reset() for b in range(nb1)]
            r1_moves = torch.randint(1, 7, (4, G1, 2), dtype=torch.uint8, device=dev, generator=gen)
            r1_res = [torch.empty((G1,), dtype=torch.uint8, device=dev) for _ in range(nb1)]
            for b in range(nb1):
                r1_batches[b].rollout(64)            # populate with enemies / items, untimed
            ms_r1, rep_r1, _, _ = timed_graph(lambda k: r1_batches[k % nb1].update(r1_moves[k % 4], out=r1_res[k % nb1]), 72, nb1)
            r1_bytes = r1_batches[0].alg_bytes_per_game_tick
            out['r1'] = {'note': 'ruleset R1 = README-only rules (docs/RULESET_R1.md); PARITY UNPINNED vs the reference, bit-exact vs '
                                 'oracle/orx_r1_oracle.c; configs[2]: 65,536 games per GPU, 8 enemy + 4 item slots',
                         'value': world * G1 / (ms_r1 * 1e-3), 'unit': UNIT, 'games_per_gpu': G1, 'us_per_step': ms_r1 * 1e3,
                         'alg_bytes_per_game_tick': r1_bytes, 'achieved': r1_bytes * G1 / (ms_r1 * 1e-3) / 1e9}
            del r1_batches
            # the same ruleset at the headline's size (SURVEY config 4 "run for R0 and R1"): this rank's shard of 2^20 games
            nb4 = max(2, -(-300_000_000 // (244 * G)))
            r1_big = [R1GameState(G, max_ticks=MAX_TICKS, auto_reset=True, seed=SEED, device=dev, overlap_ticks=True,
                                  game_id_base=(1 << 42) + (rank * nb4 + b) * G).reset() for b in range(nb4)]
            r1_big_moves = torch.randint(1, 7, (2, G, 2), dtype=torch.uint8, device=dev, generator=gen)
            r1_big_res = [torch.empty((G,), dtype=torch.uint8, device=dev) for _ in range(nb4)]
            for b in range(nb4):
                r1_big[b].rollout(64)
            ms_r1b, _, _, _ = timed_graph(lambda k: r1_big[k % nb4].update(r1_big_moves[k % 2], out=r1_big_res[k % nb4]), min(K, 24), nb4)
            out['r1']['headline_size'] = {'games_per_gpu': G, 'global_games': world * G, 'value': world * G / (ms_r1b * 1e-3), 'us_per_step': ms_r1b * 1e3,
                                          'achieved': r1_bytes * G / (ms_r1b * 1e-3) / 1e9, 'rotating_batches': nb4}
            del r1_big, r1_big_moves, r1_big_res

    if rank == 0:
        time.sleep(0.15)
        sampler.stop()
        peak, peak_src = measured_peak()
        achieved = B_ALG * G / (ms_step * 1e-3) / 1e9
        tr = recorded_traffic()
        line = {
            'metric': METRIC, 'value': world * G / (ms_step * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'replays': replays, 'window_ms': window_ms, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'int32', 'data': 'synthetic', 'config': workload_config(args, world),
            'measurement': {'l2': f'rotating {nb} independent batches per GPU ({nb * 32 * G / 1e6:.0f} MB of planes > 126 MB L2), no flush needed',
                            'launch': f'K steps captured in one CUDA graph, replayed {replays}x back to back inside one CUDA-event pair',
                            'ordering': 'throughput mode (SimConfig.overlap_ticks / ORX_PATH_TILE_FLAGS): consecutive tick launches are ordered run by run (a run = the '
                                        'tiles of one CTA), not grid by grid; every tick of a state still sees the previous one complete, run for run',
                            'numa': numa if numa is not None else 'process not bound (single NUMA node, one GPU, or topology not exposed)'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': (tr or {}).get('dram_bytes_per_launch'), 'peak_source': peak_src,
                         'kernel': 'k_step_pipe<EMPTY,CMD_BYTES>', 'alg_bytes_per_game_tick': B_ALG, 'games_per_launch': G},
            'e2e': out.pop('e2e'),
            'gpu_launches': value_launches,
        }
        if 'grid_wait_mode' in out:
            gwm = out.pop('grid_wait_mode')
            gwm['frac'] = gwm['achieved'] / peak
            line['roofline']['grid_wait_mode'] = gwm
        if 'large_batch' in out:
            lb = out.pop('large_batch')
            lb['frac'] = lb['achieved'] / peak
            line['roofline']['large_batch'] = lb
        for key in ('step_observe', 'r1'):
            if key in out:
                num = out[key].pop('hbm_frac_of', None) or out[key].get('achieved')
                out[key]['hbm_frac'] = num / peak
        if 'r1' in out and 'headline_size' in out['r1']:
            out['r1']['headline_size']['hbm_frac'] = out['r1']['headline_size']['achieved'] / peak
        line.update(out)
        line['clocks'] = sampler.summary(t_wall0, t_wall1)
        if not args.no_cpu_baseline and world == 1:     # CPU baseline: rank 0 at N=1 only (other ranks would spin on host cores)
            port = cpu_port_run_isolated(1 << 18, seconds=min(args.cpu_seconds, 4.0))
            port_obj = {'value': port['value'], 'unit': UNIT, 'cores': port['cores'], 'kind': 'port',
                        'sample': f'{port["ticks"]} game-ticks in {port["elapsed"]:.1f} s: 2^18 games, same config, oracle/orx_oracle.c oro_rollout (RandomBot x2), OpenMP over all host cores'}
            if reference_available():
                r = cpu_reference_run_isolated(reference_sample_games(G_total), seconds=args.cpu_seconds)
                ref_obj = {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'reference',
                           'sample': f'{r["ticks"]} game-ticks in {r["elapsed"]:.1f} s: {r["games"]} games ticked in lockstep by the unmodified reference loop '
                                     f'(optimax_rogue/server/main.py:110-113: on_tick + Updater.update + 2 x RandomBot.move), one process per host core, '
                                     f'Python {r["python"]} numpy {r["numpy"]}'}
                line['cpu_baseline'] = dict(ref_obj, python_reference=ref_obj, port=port_obj)
            else:
                line['cpu_baseline'] = dict(port_obj, port=port_obj)
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
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
