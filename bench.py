#!/usr/bin/env python
"""bench.py -- the audiogan GAN training step on B200 (BASELINE.json metric / config).

A "step" is one core training step: 1 discriminator update (G forward detached,
D forward on real and on fake, D backward, per-tensor clip, RMSprop) + 1 generator update (G forward,
D forward, D data-gradient, G backward incl. BPTT, clip, RMSprop) on one synthetic minibatch.
Workload (configs[1]): default generator / discriminator, per-GPU batch 64, 2 s synthetic 8 kHz
waveforms (L = 16000); N GPUs = batch-sharded data parallel (global batch 64 N, weak scaling) with a
bucketed NCCL all-reduce of each net's gradients.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the CPU oracle port of the reference step on the host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

RATE = 8000
FRAME = 200


def flops_per_sample(L, gstate=1024, dstate=1024, g_struct=None, d_struct=None, frame=200, embed=100, noise=100):
    """Algorithmic FLOPs of one core step per sample: 4 F_G + 8 F_D forward MACs, 2 FLOP/MAC.  Default nets:
    F_G = 60 475.64 L, F_D = 140 736 L (the SURVEY's figures, reproduced by the general formulas below)."""
    g_struct = g_struct or [[17, 8, 128, 16], [9, 4, 64, 32], [9, 4, 64, 32], [9, 4, 32, 32]]
    d_struct = d_struct or [[7, 2, 16], [7, 2, 32], [7, 2, 64], [7, 2, 128], [7, 2, 256], [7, 2, 512]]
    Tg = (L + frame - 1) // frame
    F_G = Tg * ((frame + embed + noise + gstate) * 4 * gstate + gstate * frame + gstate)
    cin = 1
    for k, s, hid, out in g_struct:
        F_G += (L / s) * hid * k * cin + (L / s) * (k - 1) * hid * out
        cin += out
    F_G += 3 * cin * L
    F_D, cin, T = 0.0, 1, L
    for k, s, cout in d_struct:
        T = (T + s - 1) // s
        F_D += T * cout * k * cin
        cin = cout
    Hd = dstate // 2
    F_D += T * (2 * (cin + embed) * 4 * Hd + 2 * Hd * 4 * Hd + 2 * dstate * dstate + dstate * (dstate // 2) + dstate // 2)
    return 2.0 * (4 * F_G + 8 * F_D)


D_STRUCT_SCALED = [[7, 2, 32], [7, 1, 32], [7, 2, 64], [7, 1, 64], [7, 2, 128], [7, 1, 128], [7, 2, 256], [7, 1, 256],
                   [7, 2, 512], [7, 1, 512], [7, 2, 1024], [7, 1, 1024]]     # configs[4]: 2x channels and depth (SURVEY 8(d) cfg 5)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tc_burst=1590.0, tc=1400.0, src="fallback (B200_PROFILING.md)")


# ----------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        get_r = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = get_r(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------- our arm
class KernelTimer:
    """CUDA-event pairs around every C-ABI launch (on the launching stream), grouped per kernel family."""

    def __init__(self, torch, shapes=False):
        self.torch, self.recs, self.shapes = torch, [], shapes

    def hook(self, name, args):
        tc = self.torch.cuda
        flops = 0.0
        if name.startswith("ag_gemm"):
            d = args[0]._obj
            flops = 2.0 * d.M * d.N * d.K
        elif name.startswith("ag_lstm"):
            d = args[0]._obj
            if name.endswith("fwd"):
                per = d.ndir * 4 * d.H * (d.H + d.F) + (d.F + 1) * d.H * (1 if d.F else 0)
            else:
                per = d.ndir * d.H * (4 * d.H + (d.F + 1 if d.F else 0)) + d.F * 4 * d.H
            flops = 2.0 * d.B * d.T * per
            if self.shapes:
                name = name + " H%d ndir%d F%d T%d" % (d.H, d.ndir, d.F, d.T)
        e0, e1 = tc.Event(enable_timing=True), tc.Event(enable_timing=True)
        e0.record()
        if name.startswith("ag_gemm"):
            name = name + " M%d N%d K%d" % (d.M, d.N, d.K) if self.shapes else name
        rec = [name, flops, e0, e1]
        self.recs.append(rec)
        return e1.record

    def summary(self):
        fam = {}
        for name, flops, e0, e1 in self.recs:
            ms = e0.elapsed_time(e1)
            f = fam.setdefault(name, [0.0, 0.0, 0])
            f[0] += ms
            f[1] += flops
            f[2] += 1
        return fam


def make_batches(torch, B, L, rank, nb, pinned):
    from audiogan_b200.synthetic import step_inputs
    out = []
    for i in range(nb):
        inp = step_inputs(B, L, seed=1234 + 1000 * rank + i, full_length=True)
        if pinned:
            inp = {k: v.pin_memory() for k, v in inp.items()}
        out.append(inp)
    return out


def run_ours(args):
    import torch
    import audiogan_b200 as ag
    from audiogan_b200 import dist as agd
    from audiogan_b200 import _abi

    rank, world, local = agd.init()
    if world != args.gpus:
        if args.gpus != 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks (WORLD_SIZE=%d)" % (args.gpus, args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(1234)
    B, L = args.batch, args.samples
    d_struct = D_STRUCT_SCALED if args.d_struct == "scaled" else None
    g = ag.pin_stopper(ag.Generator(embed_size=100, state_size=args.gstate)).to(dev)
    d = ag.Discriminator(embed_size=100, state_size=args.dstate, **({"cnn_struct": d_struct} if d_struct else {})).to(dev)
    fps = flops_per_sample(L, args.gstate, args.dstate, None, d_struct)
    nets = "default nets" if (args.gstate, args.dstate, args.d_struct) == (1024, 1024, "default") else (
        "generator state %d, discriminator state %d, %s discriminator conv stack" % (args.gstate, args.dstate, args.d_struct))
    g.set_mode(args.mode)
    d.set_mode(args.mode)
    agd.broadcast_parameters([g, d])
    opt_d = ag.FusedRMSprop(d.parameters(), lr=1e-4)
    opt_g = ag.FusedRMSprop(g.parameters(), lr=1e-4)
    # data-parallel gradient all-reduce: the library's peer-memory kernels over NVLink (default), or bucketed NCCL
    # (AUDIOGAN_DP=nccl; AUDIOGAN_DP_BUCKETS buckets per net)
    sync, dp_note = None, "single GPU"
    if world > 1 and os.environ.get("AUDIOGAN_DP", "peer") == "peer":
        try:
            sync = agd.PeerGradSync([g, d])
            dp_note = "two-shot all-reduce over NVLink peer memory (csrc/peer.cu), captured in the step's graph"
        except Exception as e:                                   # noqa: BLE001 -- reported, never silent
            sys.stderr.write("[bench] peer-memory all-reduce unavailable (%r): using NCCL\n" % (e,))
            dp_note = "NCCL (peer-memory path unavailable: %s)" % (str(e).splitlines()[0][:120],)
    if world > 1 and sync is None:
        nb = int(os.environ.get("AUDIOGAN_DP_BUCKETS", "1"))
        sync = agd.GradSync(nbuckets=nb)
        dp_note = dp_note if dp_note.startswith("NCCL (") else "NCCL all-reduce, %d bucket(s) per net, after backward" % nb
    # AUDIOGAN_DP_EARLY=1: packed gradient regions are all-reduced while backward is still running.  Off by default: measured
    # 0.4 ms/step SLOWER at 2 GPUs (15.65-15.72 vs 15.26-15.31 ms) -- the NCCL CTAs take SMs the recurrent kernels' clusters need
    if sync is not None and os.environ.get("AUDIOGAN_DP_EARLY", "0") == "1":
        sync.attach(g, d)

    host = make_batches(torch, B, L, rank, 2, pinned=True)
    # the per-sample lengths are host metadata (the reference carries them as numpy arrays and reads them on the host inside
    # Discriminator.forward, audiogan.py:516): they stay CPU tensors, so no device-to-host read stalls the launch queue
    on_dev = lambda h, **kw: {k: (v if k.endswith("_len") else v.to(dev, **kw)) for k, v in h.items()}
    resident = [on_dev(h) for h in host]
    h2d_bytes = sum(v.numel() * v.element_size() for k, v in host[0].items() if not k.endswith("_len"))

    def step(di):
        # train.core_step = d_update + g_update with the step's two generator forward passes run as one batched pass
        # (AUDIOGAN_SPLIT_G=1: the literal call-by-call sequence)
        di = dict(di)
        di["u_stop"] = None
        if os.environ.get("AUDIOGAN_SPLIT_G", "0") != "1":
            return ag.core_step(g, d, opt_d, opt_g, di, clip_d=1.0, clip_g=0.1, grad_sync=sync)
        m1 = ag.d_update(g, d, opt_d, di, clip=1.0, grad_sync=sync)
        gb = {"c_g": di["g_c_g"], "c_d": di["g_c_d"], "z": di["g_z"], "noise_fake": di["g_noise_fake"], "u_stop": None}
        m2 = ag.g_update(g, d, opt_g, gb, clip=0.1, grad_sync=sync)
        return m1, m2

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t)
        return ms

    for i in range(args.warmup):
        step(resident[i % 2])
    # ---- the step as ONE CUDA graph (audiogan_b200/graph.py): both updates, their backward passes, the optimizer launches and
    # the NCCL all-reduces are captured once; a timed step = copy the batch into the graph's static inputs + one graph launch.
    # --no-graph / AUDIOGAN_GRAPH=0 times the eager launch sequence instead (same kernels, ~410 launches per step from Python).
    gs, graph_note = None, "eager launches (--no-graph)"
    if args.graph:
        try:
            gs = ag.GraphedStep(g, d, opt_d, opt_g, resident[0], clip_d=1.0, clip_g=0.1, grad_sync=sync, warmup=1)
            graph_note = "%s per step (%d library launches + torch fills/copies captured)" % ("one CUDA graph" if len(gs.graphs) == 1 else "%d CUDA graph segments cut at the gradient all-reduces" % len(gs.graphs), gs.launches)
        except Exception as e:                                   # noqa: BLE001 -- reported in the JSON line, never silent
            gs, graph_note = None, "eager launches: graph capture failed: %s" % (str(e).splitlines()[0][:200],)
            sys.stderr.write("[bench] CUDA graph capture failed, timing eager launches: %r\n" % (e,))
            torch.cuda.synchronize()

    def fast_step(di):
        if gs is None:
            return step(di)
        gs.load(di)
        out = gs.replay()
        return ({"loss_d": out["loss_d"], "loss_g": out["loss_g"]}, {"loss": out["loss"]})

    for i in range(args.warmup):
        fast_step(resident[i % 2])
    # ---- headline: inputs resident in HBM
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _abi.launches
    ms = timed(lambda i: fast_step(resident[i % 2]), args.steps)
    launches = (_abi.launches - l0) / args.steps if gs is None else float(gs.launches)
    clocks = sampler.stop()
    ms_step = ms / args.steps
    audio_s = world * B * L / RATE
    value = audio_s / (ms_step * 1e-3)

    if args.quick:
        if rank == 0:
            emit({"quick": True, "ms_per_step": round(ms_step, 4), "value": round(value, 2), "gpu_launches": launches,
                  "launch_mode": graph_note, "dp_allreduce": dp_note, "n_gpus": world, "workload": "%s, per-GPU batch %d, L=%d" % (nets, B, L),
                  "step_tflops": round(world * B * fps / (ms_step * 1e-3) / 1e12, 3),
                  "recurrent_paths": dict(g._plan.last_path, **d._plan.last_path)})
        return
    # ---- e2e: host buffers, H2D of the step's inputs and D2H of the losses inside the timed region
    d2h = [0]

    # The input pipeline a training loop would run: step i+1's inputs are copied from pinned host memory on a copy stream
    # while step i computes (every step's H2D copy happens inside the timed region), and step i's three losses come back
    # through a ring of pinned buffers that the host reads three steps later (an immediate .cpu() would drain the launch queue
    # every step; with several ranks a one-step window lets host jitter on one rank stall every rank at the next collective).
    copy_stream = torch.cuda.Stream()
    NOUT = 4                    # losses are read NOUT-1 steps late: host jitter on one rank does not stall the others' queues
    pinned_out = [torch.empty(3, pin_memory=True) for _ in range(NOUT)]
    out_ev = [None] * NOUT
    losses_seen = []

    # Two device staging sets, allocated once: step i+1's H2D copies land in set (i+1) % 2 on the copy stream while step i
    # computes; the step then moves them into the graph's static inputs (a 20 MB device-to-device copy) and replays.  Nothing is
    # allocated inside the timed region.
    dkeys = [k for k in host[0] if not k.endswith("_len")]
    stage = [on_dev(host[0]), on_dev(host[1])]
    consumed = [None, None]                 # event: the main stream has finished reading staging set k

    def prefetch(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            if consumed[k] is not None:
                copy_stream.wait_event(consumed[k])
            for key in dkeys:
                stage[k][key].copy_(host[k][key], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return k, ev

    pending = [prefetch(0)]

    e2e_count = [0]

    def e2e_step(_):
        i = e2e_count[0]
        e2e_count[0] += 1
        k, ev = pending.pop()
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        pending.append(prefetch(i + 1))
        if gs is not None:
            gs.load(stage[k])
            consumed[k] = torch.cuda.Event()
            consumed[k].record(cur)
            out = gs.replay()
            res = out["losses"]
        else:
            m1, m2 = step(stage[k])
            consumed[k] = torch.cuda.Event()
            consumed[k].record(cur)
            res = torch.stack([m1["loss_d"], m1["loss_g"], m2["loss"]])
        j = (i + 1) % NOUT
        if out_ev[j] is not None:               # the losses of step i - (NOUT - 1) have landed: read them on the host
            out_ev[j].synchronize()
            losses_seen.append(float(pinned_out[j][0]))
        pinned_out[i % NOUT].copy_(res, non_blocking=True)
        e = torch.cuda.Event()
        e.record()
        out_ev[i % NOUT] = e
        d2h[0] = res.numel() * res.element_size()

    # warm-up of THIS path too: its copy-stream allocations (inputs kept alive across streams) reach their steady state
    # after a few steps; before that the caching allocator still calls cudaMalloc inside the loop
    for _ in range(max(3, args.warmup)):
        e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps) / args.steps
    assert all(v == v for v in losses_seen), "NaN loss in the end-to-end run"

    # ---- per-kernel attribution with CUDA events around every launch (same steps, instrumented)
    # (eager launches, every kernel alone on the GPU: the side-stream "shadow" scheduling of engine.py is switched off here so that
    # an event pair brackets one kernel and not whatever runs beside it)
    from audiogan_b200 import engine as _engine
    _ovl, _engine._OVERLAP = _engine._OVERLAP, False
    step(resident[0])              # untimed: torch.cuda.graph() emptied the allocator's cache, the first eager step re-mallocs
    kt = KernelTimer(torch, shapes=args.shapes)
    _abi.set_hook(kt.hook)
    ms_inst = timed(lambda i: step(resident[i % 2]), args.steps) / args.steps
    _abi.set_hook(None)
    _engine._OVERLAP = _ovl
    fam = kt.summary()
    pk = peaks()
    tot_kernel_ms = sum(v[0] for v in fam.values())
    top = max(fam.items(), key=lambda kv: kv[1][0])
    tname, (tms, tflops, tcount) = top
    achieved = (tflops / (tms * 1e-3)) / 1e12 if tms > 0 else 0.0
    # DRAM bytes per launch of that family, from the committed ncu launch list of this same command
    # (tools/ncu_summary.py traffic -> profiles/r2_traffic.json, bf16 mode); null when the capture is missing
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json" if args.mode == "bf16" else "r1_fp32_traffic.json")) as f:
            traffic = round(json.load(f)[tname.split(" ")[0]]["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"kernel": tname, "bound": "tensor", "achieved": round(achieved, 3), "peak": pk["tc"], "unit": "TFLOP/s",
                "frac": round(achieved / pk["tc"], 5), "traffic": traffic, "avg_launch_ms": round(tms / tcount, 5),
                "launches_per_step": tcount / args.steps, "share_of_kernel_time": round(tms / tot_kernel_ms, 4),
                "peak_source": pk["src"] + " (bf16_tflops_sustained, of measured)"}
    families = {k: {"ms_per_step": round(v[0] / args.steps, 4), "tflops": round((v[1] / (v[0] * 1e-3)) / 1e12, 3) if v[0] > 0 else 0,
                    "launches_per_step": v[2] / args.steps} for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])}

    out = {
        "metric": "audio_seconds_per_s", "value": round(value, 2), "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else "bf16", "data": "synthetic",
        "steps_per_s": round(1e3 / ms_step, 4),
        "step_tflops": round(world * B * fps / (ms_step * 1e-3) / 1e12, 3),
        "frac_tc_peak_whole_step": round(B * fps / (ms_step * 1e-3) / 1e12 / pk["tc"], 5),
        "recurrent_paths": dict(g._plan.last_path, **d._plan.last_path),
        "config": {"workload": "audiogan core GAN step (1 D-update + 1 G-update), %s, per-GPU batch %d, "
                               "%.1f s synthetic 8 kHz waveforms (L=%d)" % (nets, B, L / RATE, L),
                   "global_batch": world * B, "samples": L, "mode": args.mode, "parallelism": "dp%d" % world,
                   "l2_policy": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; two input batches alternate"},
        "e2e": {"value": round(audio_s / (ms_e2e * 1e-3), 2), "unit": "audio-s/s", "ms_per_step": round(ms_e2e, 4),
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h[0],
                "pipeline": "inputs double-buffered on a copy stream, losses read back three steps late"},
        "gpu_launches": launches,
        "launch_mode": graph_note,
        "dp_allreduce": dp_note,
        "clocks": clocks,
        "roofline": roofline,
        "ms_per_step_instrumented": round(ms_inst, 4),
        "kernel_families": families,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_leg_subprocess(args, steps=2, warmup=1, batch=args.cpu_batch)
    if rank == 0:
        emit(out)
    if world > 1:
        torch.distributed.destroy_process_group()


# ----------------------------------------------------------------------------------- CPU arm
def cpu_leg_subprocess(args, steps, warmup, batch):
    """cpu_step_throughput in a child process that cannot see the GPU: the reference wraps its sub-modules in NN.DataParallel
    (audiogan.py:379-410), which on a box with visible GPUs moves the inputs to cuda:0 -- its CPU path needs CUDA hidden."""
    import subprocess
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg", "--steps", str(steps), "--warmup", str(warmup),
           "--cpu-batch", str(batch), "--batch", str(args.batch), "--samples", str(args.samples)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError("CPU baseline leg failed: %s" % (r.stderr.strip().splitlines()[-1:] or ["no output"])[0])
    return json.loads(lines[-1])


def cpu_step_throughput(args, steps, warmup, batch):
    """The reference's CPU training step on the host cores, on `batch` samples of the workload's minibatch.

    kind "reference": the reference's OWN Generator / Discriminator / helper definitions (audiogan.py:1-552, located by
    oracle/ref_step.py: /root/reference in the build container, the build-time copy under oracle/_ref/ on the GPU box) driven
    through the core step's call order on stock torch CPU kernels.  kind "port" (only when neither is present): the oracle's
    functional restatement oracle/restated.py (pinned to the reference classes by tests/test_oracle.py)."""
    import torch
    from oracle import restated as O
    from oracle import ref_step as RS
    from audiogan_b200.synthetic import step_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    L = args.samples
    inp = step_inputs(batch, L, seed=1234, full_length=True)
    if RS.locate() is not None:
        t, _ = RS.time_steps(inp, steps, warmup)
        kind, what = "reference", "the reference's own classes (audiogan.py:1-552) through the core step's call order"
    else:
        Pg = O.pin_stopper(O.init_generator(11))
        Pd = O.init_discriminator(12)
        gb = {"c_g": inp["g_c_g"], "c_d": inp["g_c_d"], "z": inp["g_z"], "noise_fake": inp["g_noise_fake"]}
        sd, sg, times = {}, {}, []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.d_update(Pg, Pd, sd, inp)
            O.g_update(Pg, Pd, sg, gb)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        t = sum(times) / len(times)
        kind, what = "port", "oracle port of the reference step (oracle/restated.py)"
    return {"value": round(batch * L / RATE / t, 3), "unit": "audio-s/s", "cores": cores, "kind": kind,
            "s_per_step": round(t, 4), "steps_per_s_at_sample_batch": round(1.0 / t, 4), "sample_batch": batch,
            "sample": "%s, torch fp32 CPU, %d threads, on %d of the %d samples of a minibatch, L=%d, %d timed steps after %d "
                      "warm-up" % (what, cores, batch, args.batch, L, steps, warmup)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the step on the SAME config (full per-GPU minibatch, same L);
    each step is the whole minibatch, a bounded number of steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 2))
    warm = 1 if args.warmup > 0 else 0
    cb = cpu_leg_subprocess(args, steps=steps, warmup=warm, batch=args.batch)
    L = args.samples
    out = {
        "impl": "reference", "metric": "audio_seconds_per_s", "value": cb["value"], "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": round(cb["s_per_step"] * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "audiogan core GAN step (1 D-update + 1 G-update), default nets, per-GPU batch %d, "
                               "%.1f s synthetic 8 kHz waveforms (L=%d)" % (args.batch, L / RATE, L),
                   "global_batch": args.batch, "samples": L, "mode": "fp32", "parallelism": "cpu"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line on the real stdout (libraries such as NCCL print banners to stdout: fd 1 is pointed at stderr
    for the duration of the run)."""
    line = json.dumps(obj) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line.encode())
    else:
        sys.stdout.write(line)
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("AUDIOGAN_MODE", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch (configs[1]: 64)")
    ap.add_argument("--samples", type=int, default=16000, help="waveform length L (2 s at 8 kHz)")
    ap.add_argument("--gstate", type=int, default=1024, help="generator state size (configs[3]: 2048)")
    ap.add_argument("--dstate", type=int, default=1024, help="discriminator state size (configs[3]: 2048)")
    ap.add_argument("--d-struct", default="default", choices=["default", "scaled"],
                    help="scaled = configs[4]: 2x conv channels and depth (added layers stride 1)")
    ap.add_argument("--cpu-batch", type=int, default=8, help="samples per step of the bounded CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shapes", action="store_true", help="attribute GEMM time per (M,N,K) shape")
    ap.add_argument("--quick", action="store_true", help="headline timing only (for ncu runs): no e2e / attribution / CPU legs")
    ap.add_argument("--no-graph", dest="graph", action="store_false", default=os.environ.get("AUDIOGAN_GRAPH", "1") != "0",
                    help="time the eager launch sequence instead of the captured CUDA graph")
    ap.add_argument("--cpu-leg", action="store_true", help=argparse.SUPPRESS)     # child process of cpu_leg_subprocess
    args = ap.parse_args()
    if args.cpu_leg:
        emit(cpu_step_throughput(args, steps=args.steps, warmup=args.warmup, batch=args.cpu_batch))
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
