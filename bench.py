#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: path-steps/s of the fused Monte Carlo kernels on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload merton|gbm|...] [--paths P]
                    [--scaling weak|strong] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic input: simulate `paths` paths for `num_steps` time
steps, apply the payoff and reduce (sum, sum^2).  Default workload = the configuration BASELINE.json's north_star
target is quoted on: Merton 1-D jump-diffusion European call, jump-adapted Euler, 1e9 paths x 100 steps per GPU,
payoff at the last state ('adapted'; mu=.02 sigma=.2 rate=1 alpha=-.05 gamma=.3 S0=K=1 T=3, the parameters of
examples/merton_1d_european).  `--workload gbm` is BASELINE.json configs[1] (GBM 1-D, Euler-Maruyama, 1e9 x 252).

Scaling.  weak (default): every rank simulates `paths` paths of a disjoint global path-id range.  strong: `paths` is
the TOTAL, split evenly over the ranks ("N = 1e9 split evenly over G"); because the Philox counter is
the global path id the estimate must not depend on G -- `estimate.mean_repr` prints it to 17 digits.  Either way the
only exchange is one all-reduce of 8 fp64 moments.  Prints ONE JSON line on rank 0.

`--impl reference` times the reference's own CPU implementation on the host cores, on a bounded sample of the same
workload: the UNMODIFIED reference package from oracle/_ref (installed there by `make -C oracle ref`; "kind":
"reference") or, if that copy did not travel, oracle/torch_port.py ("kind": "port", pinned bit-for-bit to it).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Native libraries write to file descriptor 1 behind Python's back (NCCL prints
# its version banner there when the first communicator is created), so fd 1 is pointed at stderr for the whole run
# and the result line goes to the saved original descriptor.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_RESULT_FD, (line + "\n").encode())

# canonical algorithmic work per nominal path-step, in FP32/ALU/XU lane-operations (SURVEY.md section 8d,
# restated in DESIGN.md): one N(0,1) = 21 ops; GBM Euler step 2 ops; Merton jump-adapted iteration 39 ops x 1.03.
OPS_PER_PATH_STEP = {"gbm": 23.0, "merton": 40.0, "levy2d": 240.0, "merton_cv": 40.0}
WORKLOADS = {
    # configs[0] at the size the north-star target is quoted on (the default)
    "merton": dict(name="merton_1d_eurocall_jump_adapted_euler_1e9x100", num_steps=100, paths=10 ** 9,
                   cpu_paths=10 ** 5),
    # BASELINE.json configs[1]
    "gbm": dict(name="gbm_1d_eurocall_euler_1e9x252", num_steps=252, paths=10 ** 9, cpu_paths=10 ** 5),
    # configs[3]: 2-D Levy-driven rainbow, 1e8 paths x 256 steps
    "levy2d": dict(name="levy_2d_rainbow_jump_adapted_euler_1e8x256", num_steps=256, paths=10 ** 8, cpu_paths=5000),
    # configs[2]: Merton 1-D with the neural control variate applied in-kernel (tcgen05), 1e8 paths x 200 steps
    "merton_cv": dict(name="merton_1d_neural_cv_tcgen05_1e8x200", num_steps=200, paths=10 ** 8, cpu_paths=10 ** 4),
}
WORKLOADS.update({
    # path-storing mode (the solve() contract): HBM-write-bound; bytes per path = reference layouts
    "gbm_store": dict(name="gbm_1d_solve_store_paths_4e6x252", num_steps=252, paths=4 * 10 ** 6, cpu_paths=10 ** 5),
    "merton_store": dict(name="merton_1d_solve_full_storage_2e6x100", num_steps=100, paths=2 * 10 ** 6,
                         cpu_paths=10 ** 5),
})
WORKLOADS.update({
    # configs[4]: MLMC Merton 1-D, coupled fine/coarse levels 1..128, allocation of mlmc.py:77-97 at eps = 1e-4
    "mlmc": dict(name="merton_1d_mlmc_levels_1_2_4_to_128_eps1e-4", num_steps=1, paths=1, cpu_paths=1),
})
DEFAULT_WORKLOAD = "merton"
MLMC_LEVELS = [1, 2, 4, 8, 16, 32, 64, 128]
MLMC_EPS = 1e-4
OPS_PER_PATH_STEP.update({"gbm_store": 23.0, "merton_store": 40.0, "mlmc": 40.0})
# What ncu saw of each workload's dominant kernel.  NOT measured in this run (a bench number is never taken under a
# profiler): copied from the committed `ncu --set full` captures named in `source`, taken with the same library on the
# same workload at the path count `capture_paths`; the JSON line repeats the source next to every figure derived
# from them.
#   dram_bytes  : dram__bytes_read.sum + dram__bytes_write.sum of one launch of the capture
#   inst        : executed thread-instructions per path-ITERATION = smsp__inst_executed.sum x 32 / (paths x
#                 iterations per path) -- the executed counterpart of the canonical OPS_PER_PATH_STEP
#   binding     : the busiest pipe of the capture and its thread-instructions per path-iteration and lanes per SM
#   pipe_pct    : pipe utilisation in the capture (percent of peak while active)
PROFILE = {
    "gbm": dict(source="profiles/r02_ncu_gbm.summary.txt", capture_paths=1e8, dram_bytes=20224.0, inst=14.59,
                binding=dict(pipe="xu", inst=2.02, lanes_per_sm=16),
                pipe_pct=dict(issue=65.5, fma=32.0, alu=44.9, xu=72.3)),
    "merton": dict(source="profiles/r02_ncu_merton.summary.txt", capture_paths=5e7, dram_bytes=45056.0, inst=37.36,
                   binding=dict(pipe="issue", inst=37.36, lanes_per_sm=128),
                   pipe_pct=dict(issue=69.1, fma=34.0, alu=39.9, xu=45.7)),
    "levy2d": dict(source="profiles/r02_ncu_levy2d.summary.txt", capture_paths=5e6, dram_bytes=36608.0, inst=145.3,
                   binding=dict(pipe="issue", inst=145.3, lanes_per_sm=128),
                   pipe_pct=dict(issue=65.6, fma=31.7, alu=45.9, xu=46.2)),
    "merton_cv": dict(source="profiles/r02_ncu_merton_cv.summary.txt", capture_paths=2e6, dram_bytes=189440.0,
                      inst=667.7, binding=None, pipe_pct=dict(issue=43.6, fma=5.6, alu=44.9, xu=6.2, tensor=41.7)),
    "mlmc": dict(source="profiles/r02_ncu_mlmc.summary.txt", capture_paths=None, dram_bytes=15616.0, inst=None,
                 binding=None, pipe_pct=dict(issue=67.6, fma=25.8, alu=54.4, xu=43.5)),
    "gbm_store": dict(source="profiles/r02_ncu_gbm_store.summary.txt", capture_paths=4e6, dram_bytes=8.13366e9,
                      inst=None, binding=None, pipe_pct=dict(issue=57.1, fma=20.9, alu=33.4, xu=34.6)),
    "merton_store": dict(source="profiles/r02_ncu_merton_store.summary.txt", capture_paths=2e6, dram_bytes=5.40505e9,
                         inst=None, binding=None, pipe_pct=dict(issue=41.0, fma=16.1, alu=25.4, xu=15.1)),
}
CV_TENSOR_FLOP_PER_ITER = 20000.0  # 2 nets x 2 hidden layers x 2*50*50 (unpadded)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--paths", type=float, default=None,
                    help="paths per step: per GPU (weak scaling) or in total (strong); default: the config's 1e9")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def build_problem(sm, workload, device):
    import torch
    if workload in ("gbm", "gbm_store"):
        p = sm.BlackScholesEuroCall.default_params(252, device)
        return p.solver, p.payoff, p.discounter, "terminal", None
    if workload == "levy2d":
        levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
        sde = sm.LevySde(levy, torch.tensor([1., 1.]))
        return sm.JumpEulerSolver(sde, 3, 256, device=device), sm.Rainbow(1.0), sm.ConstantShortRate(0.02), "adapted", None
    sde = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    if workload == "mlmc":
        # exact_jumps=True: the coupled pair then telescopes exactly (with the reference's default
        # the level corrections carry a -1.4e-3 bias, larger than the 1e-4 target)
        return (sm.JumpEulerSolver(sde, 3, 1, device=device, exact_jumps=True), sm.EuroCall(1.0),
                sm.ConstantShortRate(0.02), "adapted", None)
    if workload == "merton_cv":
        # the experiments' architecture (merton_cv_experiment.py:37-38), TRAINED by the stock pipeline before anything
        # is timed (:41-45: 1e4 stored jump-adapted paths of a 100-step grid, 10 epochs of Adam), so the benchmarked
        # control variates actually reduce variance; there are no checkpoints offline
        torch.manual_seed(0)
        nets = [sm.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False, device=device) for _ in range(2)]
        solver = sm.JumpEulerSolver(sde, 3, 100, device=device)
        call, csr = sm.EuroCall(1.0), sm.ConstantShortRate(0.02)
        adam = torch.optim.Adam([w for n in nets for w in n.parameters()])
        dl = sm.simulate_adapted_data(10 ** 4, solver, call, csr, bs=1000)
        sm.train_adapted_control_variates(nets, adam, dl, solver, csr, 10, False)
        for n in nets:
            n.eval()
        solver.num_steps = 200
        return solver, call, csr, "adapted", nets
    return (sm.JumpEulerSolver(sde, 3, 100, device=device), sm.EuroCall(1.0), sm.ConstantShortRate(0.02), "adapted", None)


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU with nvidia-smi while the timed region runs"""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm_mhz, max_mhz, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm_mhz.append(float(s[0]))
                max_mhz = max(max_mhz, float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm_mhz) if sm_mhz else None, "sm_max_mhz": max_mhz or None,
                "reasons": sorted(reasons), "samples": len(sm_mhz)}


def _import_reference():
    """the UNMODIFIED reference package from oracle/_ref (test infrastructure, see oracle/Makefile: `make ref`)"""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "sde_mc")):
        return None
    sys.path.insert(0, ref_dir)
    try:
        import sde_mc
        return sde_mc
    except Exception as e:              # a broken copy must not take the bench down: the port is pinned to it
        print("oracle/_ref not importable (%s); timing the port" % e, file=sys.stderr)
        return None
    finally:
        sys.path.remove(ref_dir)


def _reference_runner(ref, workload, n):
    """one pass of `workload` over n paths through the reference's own public API (mc.py / mlmc.py / solvers.py);
    returns (callable -> (mean, stderr), path-steps per call, description)"""
    import torch
    w = WORKLOADS[workload]
    merton = lambda: ref.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1)
    call, csr = ref.EuroCall(1.0), ref.ConstantShortRate(0.02)
    stat = lambda s: (float(s.sample_mean), float(s.sample_std))
    if workload == "mlmc":
        # the reference's MLMC only runs in fp64 (its fp32 pair asserts, solvers.py:264).  Bounded sample: every level
        # of the eps = 1e-4 allocation at 1/2000 of its size, coupled pairs included
        shape = [3.7e8, 4.2e7, 2.6e7, 1.5e7, 7.8e6, 4.1e6, 2.1e6, 1.0e6]
        counts = [max(int(c / 2000), 64) for c in shape]
        work = sum(c * l for c, l in zip(counts, MLMC_LEVELS))

        def run():
            torch.set_default_dtype(torch.float64)
            try:
                solver = ref.JumpEulerSolver(merton(), 3, 1, exact_jumps=True)
                return stat(ref.mc_multilevel(counts, MLMC_LEVELS, solver, call, csr))
            finally:
                torch.set_default_dtype(torch.float32)
        return run, work, "mc_multilevel, levels 1..128 at 1/2000 of the eps=1e-4 allocation, fp64"
    if workload == "gbm_store":
        solver = ref.EulerSolver(ref.Gbm(0.02, 0.3, torch.tensor([1.]), 1), 3, 252)
        return (lambda: (float(solver.solve(bs=n)[0][:, -1].mean()), 0.0)), n * 252, "EulerSolver.solve(bs=%d)" % n
    if workload == "merton_store":
        solver = ref.JumpEulerSolver(merton(), 3, 100)
        return (lambda: (float(solver.solve(bs=n)[0][:, -1].mean()), 0.0)), n * 100, "JumpEulerSolver.solve(bs=%d)" % n
    if workload == "gbm":
        p = ref.BlackScholesEuroCall.default_params(252, 'cpu')
        return (lambda: stat(ref.mc_simple(n, p.solver, p.payoff, p.discounter, bs=n))), n * 252, \
            "mc_simple(%d, EulerSolver(Gbm, 3, 252), EuroCall, bs=%d)" % (n, n)
    if workload == "levy2d":
        levy = ref.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
        solver = ref.JumpEulerSolver(ref.LevySde(levy, torch.tensor([1., 1.])), 3, 256)
        return (lambda: stat(ref.mc_simple(n, solver, ref.Rainbow(1.0), csr, bs=n, payoff_time='adapted'))), n * 256, \
            "mc_simple(%d, JumpEulerSolver(LevySde, 3, 256), Rainbow, bs=%d, 'adapted')" % (n, n)
    if workload == "merton_cv":
        torch.manual_seed(0)
        nets = [ref.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False).eval() for _ in range(2)]
        solver = ref.JumpEulerSolver(merton(), 3, 200)
        return (lambda: stat(ref.mc_apply_cvs(nets, solver, n, call, csr, sim_bs=n, bs=1000))), n * 200, \
            "mc_apply_cvs([f, g], JumpEulerSolver(Merton, 3, 200), %d, sim_bs=%d, bs=1000)" % (n, n)
    solver = ref.JumpEulerSolver(merton(), 3, 100)
    return (lambda: stat(ref.mc_simple(n, solver, call, csr, bs=n, payoff_time='adapted'))), n * 100, \
        "mc_simple(%d, JumpEulerSolver(Merton, 3, 100), EuroCall, bs=%d, 'adapted')" % (n, n)


def _port_runner(workload, n):
    """the same passes through oracle/torch_port.py (pinned to the reference on seeded runs, tests/test_torch_port.py)"""
    import torch
    import sde_mc_b200 as sm
    from oracle import torch_port as tp
    merton = sm.Merton(0.02, 0.2, 1, -0.05, 0.3, torch.tensor([1.]), 1).kernel_spec()
    if workload == "mlmc":
        shape = [3.7e8, 4.2e7, 2.6e7, 1.5e7, 7.8e6, 4.1e6, 2.1e6, 1.0e6]
        counts = [max(int(c / 2000), 64) for c in shape]

        def run():   # single-level simulation only: the port has no coupled pair (flatters the CPU)
            last = 0.0
            for c, l in zip(counts, MLMC_LEVELS):
                last = float(tp.jump_solve(merton, 3, l, c, low_storage=True)[0][:, -1].mean())
            return last, 0.0
        return run, sum(c * l for c, l in zip(counts, MLMC_LEVELS)), "fine paths of every level only (no coupled pair)"
    if workload == "gbm_store":
        spec = sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1).kernel_spec()
        return (lambda: (float(tp.diffusion_solve(spec, 3, 252, n)[0][:, -1].mean()), 0.0)), n * 252, "diffusion_solve"
    if workload == "merton_store":
        return (lambda: (float(tp.jump_solve(merton, 3, 100, n, low_storage=False)[0][:, -1].mean()), 0.0)), n * 100, \
            "jump_solve"
    if workload == "gbm":
        spec = sm.Gbm(0.02, 0.3, torch.tensor([1.]), 1).kernel_spec()
        return (lambda: tp.mc_simple_batched(spec, 3, 252, n, n, tp.payoff_call_on("euro_call", 1.0), 0.02, False,
                                             "terminal")[:2]), n * 252, "mc_simple_batched"
    if workload == "levy2d":
        levy = sm.ExpExampleLevy(1, 1, 0.5, 2, 0.02, 0.3, 0.2, 0.001, dim=2)
        spec = sm.LevySde(levy, torch.tensor([1., 1.])).kernel_spec()
        return (lambda: tp.mc_simple_batched(spec, 3, 256, n, n, tp.payoff_call_on("rainbow", 1.0), 0.02, True,
                                             "adapted")[:2]), n * 256, "mc_simple_batched"
    if workload == "merton_cv":
        torch.manual_seed(0)
        nets = [sm.Mlp(2, [50, 50, 50], 1, batch_norm=False, batch_norm_init=False).eval() for _ in range(2)]
        return (lambda: tp.mc_apply_cvs_batched(merton, 3, 200, n, nets, 0.02, tp.payoff_call_on("euro_call", 1.0),
                                                1000)[:2]), n * 200, "mc_apply_cvs_batched"
    return (lambda: tp.mc_simple_batched(merton, 3, 100, n, n, tp.payoff_call_on("euro_call", 1.0), 0.02, True,
                                         "adapted")[:2]), n * 100, "mc_simple_batched"


def cpu_reference(workload, steps, warmup, sample_paths=None):
    """the reference's CPU implementation on a bounded sample, all host threads; returns (path-steps/s, s/step, info)"""
    import torch
    w = WORKLOADS[workload]
    n = int(sample_paths or w["cpu_paths"])
    torch.set_num_threads(os.cpu_count() or 1)
    ref = _import_reference()
    if ref is not None:
        run, work, what = _reference_runner(ref, workload, n)
        kind, impl = "reference", "unmodified piers-hinds/sde_mc %s from oracle/_ref" % getattr(ref, "__version__", "")
    else:
        run, work, what = _port_runner(workload, n)
        kind, impl = "port", "oracle/torch_port.py"
    torch.manual_seed(1)
    for _ in range(warmup):
        run()
    times, est = [], None
    for _ in range(steps):
        t0 = time.perf_counter()
        est = run()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = work * steps / total
    info = {"value": value, "unit": "path-steps/s", "cores": torch.get_num_threads(), "kind": kind,
            "sample": "%s; %d path-steps per step, %d timed steps, torch %s eager CPU (%s)" %
                      (what, work, steps, torch.__version__, impl),
            "estimate": est[0], "stderr": est[1]}
    return value, total / steps, info


def main():
    args = parse()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        k, wu = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
        value, sec, info = cpu_reference(args.workload, k, wu)
        emit(json.dumps({
            "impl": "reference", "metric": "path_steps_per_sec", "value": value, "unit": "path-steps/s",
            "n_gpus": args.gpus, "steps": k, "warmup": wu, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "sample": info["sample"]}, "cpu_baseline": info,
            "e2e": {"value": value, "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    import torch
    import torch.distributed as dist
    import sde_mc_b200 as sm
    from sde_mc_b200 import _engine as E
    from sde_mc_b200 import _lib as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.load()
    paths = int(args.paths or w["paths"])
    strong = args.scaling == "strong"
    # paths one step simulates over ALL ranks: weak = `paths` per GPU, strong = `paths` in total (split by
    # _engine.shard: contiguous global path-id ranges, so the estimate is the same for every G)
    total_paths = paths if strong else paths * world
    solver, payoff, discounter, payoff_time, nets = build_problem(sm, args.workload, dev)
    index_mode = L.INDEX_ADAPTED if payoff_time == "adapted" else L.INDEX_TERMINAL

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident leg: inputs (a parameter struct + Philox key) are already on the device side ----
    storing = args.workload.endswith("_store")
    store_bytes = [0]
    mlmc = args.workload == "mlmc"
    mlmc_trials = None
    if mlmc:
        from sde_mc_b200 import mlmc as M
        mlmc_trials = sm.get_optimal_trials(10 ** 5, MLMC_LEVELS, MLMC_EPS, solver, payoff, discounter)  # untimed pilot
        paths = sum(int(nl) * lv for nl, lv in zip(mlmc_trials, MLMC_LEVELS))   # fine path-steps per pass
        w = dict(w, num_steps=1)

    def mlmc_step():
        # all levels queued back to back (each rank 1/G of every level), moments stay on the device
        return M._all_levels(solver, payoff, discounter, [int(v) for v in mlmc_trials], MLMC_LEVELS)

    from sde_mc_b200._engine import shard
    local_paths = shard(total_paths, rank, world)[1]     # storing mode: each rank stores its own share

    def store_step():
        # the solve() contract: trajectories in the reference's layouts, resident in HBM (each rank its own paths)
        out = solver.solve(bs=local_paths)
        tensors = [out[0]] + [t for t in (out[1] if isinstance(out[1], tuple) else (out[1],)) if torch.is_tensor(t)]
        store_bytes[0] = sum(t.numel() * t.element_size() for t in tensors) if not solver.has_jumps else \
            sum(t.numel() * t.element_size() for t in tensors[1:]) + out[0].shape[0] * (solver.num_steps + solver.max_jumps + 1) * out[0].shape[2] * 4
        return out

    def device_step():
        # every rank: `paths` paths of its own global path-id range (weak scaling), then the 64-byte all-reduce
        if storing:
            return store_step()
        if mlmc:
            return mlmc_step()
        if nets is not None:
            return sm.mc_cv_fused(nets, solver, total_paths, payoff, discounter)
        return E.run_moments(solver, payoff, discounter, total_paths, index_mode)

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    device_step()          # untimed: the GPU idled while the clock sampler started; bring it back under load
    stream = torch.cuda.current_stream(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if storing:
        solver.kernel_events = []     # (start, end) events around every solve() kernel launch of the timed region
    ev0.record(stream)
    mom = None
    for _ in range(args.steps):
        if storing:
            mom = None     # a batch of trajectories is dropped before the next one is simulated (as in the e2e leg)
        mom = device_step()
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    kernel_ms = None
    if storing:
        kernel_ms = sum(a.elapsed_time(b) for a, b in solver.kernel_events) / max(len(solver.kernel_events), 1)
        solver.kernel_events = None
    mlmc_result = None
    if mlmc:
        tot_mean, tot_var, tot_iters, tot_n = 0.0, 0.0, 0.0, 0.0
        for nl, r in zip(mlmc_trials, mom.read()):      # one (levels, 8) tensor: one read
            mean_l, se_l = E.mean_and_stderr(r["sum"], r["sumsq"], int(nl))
            tot_mean += mean_l
            tot_var += se_l * se_l
            tot_iters += r["iters"]
            tot_n += r["n"]
        mlmc_result = {"mean": tot_mean, "stderr": tot_var ** 0.5}
        result = {"sum": 0.0, "sumsq": 0.0, "n": tot_n, "iters": tot_iters}
    elif storing:
        last = mom
        mom = None
        final = last[0][:, -1, 0].double()
        result = {"sum": float(final.sum()), "sumsq": float((final * final).sum()), "n": float(local_paths),
                  "iters": float(local_paths) * w["num_steps"]}
        del last, final
    else:
        result = mom.read()

    # ---- end-to-end leg: the public API call a user makes, host struct in -> python floats out ----
    barrier()
    t0 = time.perf_counter()
    stats = None
    for _ in range(args.steps):
        if storing:
            stats = store_step()
            stats = None
        elif mlmc:
            stats = sm.mc_multilevel(mlmc_trials, MLMC_LEVELS, solver, payoff, discounter)
        elif nets is not None:
            stats = sm.mc_apply_cvs(nets, solver, total_paths, payoff, discounter, sim_bs=10 ** 5, bs=2000)
        else:
            stats = sm.mc_simple(total_paths, solver, payoff, discounter, bs=10 ** 6, payoff_time=payoff_time)
    torch.cuda.synchronize(dev)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if rank == 0:
        time.sleep(0.2)
        sampler.stop()

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()

    if rank == 0:
        total_path_steps = float(total_paths if not mlmc else paths * world) * w["num_steps"] * args.steps
        value = total_path_steps / (dev_ms * 1e-3)
        e2e = total_path_steps / (e2e_ms * 1e-3)
        clocks = sampler.summary()
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = clocks["sm_mhz"] or clocks["sm_max_mhz"] or 1965.0
        peak = sm_count * 128 * mhz * 1e6 / 1e12                      # Tlane-op/s at the clock measured under load
        achieved = OPS_PER_PATH_STEP[args.workload] * (value / world) / 1e12
        n_total = total_paths
        import ctypes
        h2d_bytes = ctypes.sizeof(L.SdemcSde) + ctypes.sizeof(L.SdemcPayoff) + ctypes.sizeof(L.SdemcRange) + 40 * 2
        mean, se = E.mean_and_stderr(result["sum"], result["sumsq"], n_total)
        prof = PROFILE.get(args.workload, {})
        # executed-instruction view of the same throughput (this run's iterations/s x the capture's instruction counts)
        iters_per_s_per_gpu = result["iters"] / max(result["n"], 1.0) * (float(total_paths) * args.steps) / (dev_ms * 1e-3) / world
        frac_issue = frac_pipe = None
        if prof.get("inst"):
            frac_issue = prof["inst"] * iters_per_s_per_gpu / 1e12 / peak
        if prof.get("binding"):
            bnd = prof["binding"]
            frac_pipe = bnd["inst"] * iters_per_s_per_gpu / (sm_count * bnd["lanes_per_sm"] * mhz * 1e6)
        out = {
            "metric": "path_steps_per_sec", "value": value, "unit": "path-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "paths_per_gpu_per_step": total_paths // world,
                       "total_paths_per_step": total_paths, "num_steps": w["num_steps"],
                       "payoff_time": payoff_time, "moments": "fp64", "rng": "philox4x32-10 in registers",
                       "l2": "kernel reads no global inputs (parameters in constant bank); nothing to flush",
                       "parallelism": "paths sharded over %d GPU(s) by global path id, one 64-byte all-reduce per step" % world},
            "e2e": {"value": e2e, "unit": "path-steps/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 64,
                    "ms_per_step": e2e_ms / args.steps,
                    "call": ("sde_mc_b200.mc_apply_cvs([f, g], solver, paths, payoff, discounter)" if nets is not None
                             else "sde_mc_b200.mc_simple(paths, solver, payoff, discounter, bs=1e6, payoff_time=%r)" % payoff_time)},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "roofline": {"bound": "fp32_issue", "achieved": achieved, "peak": peak, "unit": "Tlaneop/s",
                         "frac": achieved / peak,
                         "frac_def": "canonical lane-ops per nominal path-step (SURVEY 8d) x path-steps/s / peak; the SASS "
                                     "executes fewer instructions than the canonical count, so this is a work rate, "
                                     "not a utilisation -- frac_issue is the utilisation",
                         "ops_per_path_step": OPS_PER_PATH_STEP[args.workload],
                         "peak_def": "%d SMs x 128 FP32 lanes x %.0f MHz (median SM clock sampled during the run)"
                                     % (sm_count, mhz),
                         "executed_iterations_per_path": result["iters"] / max(result["n"], 1.0),
                         "frac_issue": frac_issue,
                         "frac_binding_pipe": frac_pipe,
                         "binding_pipe": (prof.get("binding") or {}).get("pipe"),
                         "executed_inst_per_path_iteration": prof.get("inst"),
                         "traffic": prof.get("dram_bytes"),
                         "ncu_pipe_pct": prof.get("pipe_pct"),
                         "source": prof.get("source"),
                         "source_note": "traffic, ncu_pipe_pct and the instruction counts behind frac_issue / "
                                        "frac_binding_pipe come from the committed ncu capture `source` (%s paths), "
                                        "not from this run; throughput, iterations and clocks are this run's"
                                        % prof.get("capture_paths")},
            "estimate": {"mean": mean, "mean_repr": "%.17g" % mean, "stderr": se, "n": n_total,
                         "closed_form": {"gbm": 0.22943206, "levy2d": None}.get(args.workload, 0.26298121)},
        }
        if storing:
            # roofline: the kernel's own launch duration (CUDA events around the launch inside solve(), on its stream);
            # `value` / ms_per_step are the whole solve() call: allocation, launch, the host read of total_steps
            gbs = store_bytes[0] / (kernel_ms * 1e-3) / 1e9
            peaks = {}
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                    peaks = json.load(fh)
            except OSError:
                pass
            hpeak = peaks.get("hbm_gbs", 6650.0)
            out["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hpeak, "unit": "GB/s", "frac": gbs / hpeak,
                               "traffic": prof.get("dram_bytes"), "source": prof.get("source"),
                               "ncu_pipe_pct": prof.get("pipe_pct"), "bytes_per_step": store_bytes[0],
                               "kernel_ms_per_launch": kernel_ms,
                               "achieved_def": "algorithmic bytes of the reference layouts / mean duration of the "
                                               "storing kernel's launches in the timed region (CUDA events recorded "
                                               "around the launch inside solve()); ms_per_step is the whole call",
                               "peak_def": "MEASURED_PEAKS.json hbm_gbs (copy, read+write)" if peaks else "fallback 6.65 TB/s"}
            out["e2e"]["call"] = "solver.solve(bs=paths)  (trajectories stay on the device, as in the reference)"
            out["e2e"]["d2h_bytes_per_step"] = 0
            out["estimate"] = {"mean_terminal_state": result["sum"] / result["n"], "n": result["n"]}
        if mlmc:
            # world == 1: `paths` = fine path-steps of one pass; every rank simulates 1/G of each level (strong scaling)
            out["scaling"] = "strong"
            out["value"] = value / world
            out["e2e"]["value"] = e2e / world
            out["e2e"]["call"] = "sde_mc_b200.mc_multilevel(trials, levels, solver, payoff, discounter)"
            out["e2e"]["d2h_bytes_per_step"] = 64 * len(MLMC_LEVELS)     # ONE read of the (levels, 8) fp64 tensor
            out["gpu_launches"] = args.steps * len(MLMC_LEVELS)
            out["config"].update({"levels": MLMC_LEVELS, "eps": MLMC_EPS, "trials_per_level": [int(v) for v in mlmc_trials],
                                  "fine_path_steps_per_pass": paths, "exact_jumps": True,
                                  "parallelism": "every level sharded over %d GPU(s), ONE all-reduce of the (levels, 8) fp64 moments per pass" % world})
            out["roofline"]["achieved"] = OPS_PER_PATH_STEP["mlmc"] * out["value"] / 1e12 / world
            out["roofline"]["frac"] = out["roofline"]["achieved"] / peak
            # `value` counts NOMINAL fine steps (a level-0 path is one step), but a jump-adapted path executes
            # num_steps + #jumps iterations and level 0 -- half the pass -- is 1 nominal step with ~3 jumps: the same
            # 39 canonical ops per EXECUTED fine sub-step, for comparison
            exec_steps_per_s = result["iters"] / (dev_ms * 1e-3) / world   # last pass's executed sub-steps / ms per pass
            out["roofline"]["executed_fine_substeps_per_pass"] = result["iters"]
            out["roofline"]["frac_on_executed_iterations"] = 39.0 * exec_steps_per_s * args.steps / 1e12 / peak
            out["estimate"] = {"mean": mlmc_result["mean"], "stderr": mlmc_result["stderr"], "closed_form": 0.26298121,
                               "rmse_vs_closed_form": ((mlmc_result["mean"] - 0.26298121) ** 2 + mlmc_result["stderr"] ** 2) ** 0.5}
        if nets is not None:
            iters_per_s = result["iters"] / (dev_ms * 1e-3) * args.steps / world
            tf = CV_TENSOR_FLOP_PER_ITER * iters_per_s / 1e12
            peaks = {}
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                    peaks = json.load(fh)
            except OSError:
                pass
            tpeak = peaks.get("bf16_tflops_sustained", 1400.0)
            # the control slot of the CV kernel's moments holds the plain payoff: variance reduction of the trained nets
            nn_ = max(result["n"], 2.0)
            var_cv = result["sumsq"] / nn_ - (result["sum"] / nn_) ** 2
            var_plain = result["sumsq_c"] / nn_ - (result["sum_c"] / nn_) ** 2
            out["estimate"]["variance_reduction"] = var_plain / var_cv if var_cv > 0 else None
            out["estimate"]["plain_mean"] = result["sum_c"] / nn_
            out["config"]["nets"] = "Mlp(2,[50,50,50],1) x 2, trained untimed by train_adapted_control_variates (1e4 paths, 10 epochs)"
            out["roofline_tensor"] = {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
                                      "frac": tf / tpeak, "traffic": prof.get("dram_bytes"), "source": prof.get("source"),
                                      "flop_per_path_iteration": CV_TENSOR_FLOP_PER_ITER,
                                      "peak_def": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PF"}
        if world == 1 and not args.no_cpu_baseline:
            _, _, info = cpu_reference(args.workload, 2, 1)
            out["cpu_baseline"] = info
            if "mean" in out["estimate"] and info.get("stderr"):
                # SURVEY 8d "RMSE vs CPU": distance of the two estimates in units of their combined standard error
                est = out["estimate"]
                out["estimate"]["rmse_vs_cpu"] = abs(est["mean"] - info["estimate"]) / \
                    (est["stderr"] ** 2 + info["stderr"] ** 2) ** 0.5
                out["estimate"]["cpu_mean"], out["estimate"]["cpu_stderr"] = info["estimate"], info["stderr"]
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
