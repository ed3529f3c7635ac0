#!/usr/bin/env python
"""Benchmark of the Aligner hot path (log-likelihood GEMM+epilogue -> MAS -> durations).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]

One "step" = one pass of the hot path over one synthetic LJSpeech-shaped batch per GPU
(default workload cfg3 = BASELINE.json configs[2], the batch-256 ragged configuration the
north-star target is quoted on).  Rank 0 prints ONE JSON line.  Under torchrun every rank
owns its own batch (sharded by utterance, no collective on the data path): weak scaling.

Keys beyond the base contract:
  roofline      the dominant kernel of the step, timed live with CUDA events
  kernels       per-kernel time / algorithmic bytes / fraction of the measured HBM peak
  cpu_baseline  the reference's CPU route on the whole batch: torch-CPU log-likelihood ops + its own numba b_mas (oracle/_ref)
  e2e           same metric through the public API with pinned HOST buffers (H2D + D2H timed)
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned utterances/sec"
UNIT = "utterances/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 0.0))), "measured"
    return 6650.0, 1590.0, "fallback"      # /opt/skills/guides/B200_PROFILING.md


# ------------------------------------------------------------------------------------------
# algorithmic work per launch (restated in DESIGN.md)
# ------------------------------------------------------------------------------------------
def mas_bytes(text_len, mel_len, B, T1, T2):
    valid = int((text_len * mel_len).sum())
    return 4 * valid + 2 * B * T1 * T2 + 8 * B * T2 + 16 * B


def loglik_bytes(B, T1, T2, D, elem):
    return elem * D * B * (T1 + T2) + 8 * B * T1 * T2


def loglik_flops(B, T1, T2, D):
    return 2 * D * B * T1 * T2


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU route on the host cores
# ------------------------------------------------------------------------------------------
def base_config(w, tl, ml):
    return {"workload": w.name, "batch_per_gpu": w.batch, "t_text_max": w.t2max, "t_mel_max": w.t1max, "attention_dim": w.dim,
            "ragged": w.ragged, "valid_cells_per_batch": int((tl * ml).sum()), "padded_cells_per_batch": w.batch * w.t1max * w.t2max}


class CpuPath:
    """The hot path as the reference runs it on a CPU (alignment.py:189-208, 305-312, 275): the log-likelihood op sequence in
    torch (the reference has no separately callable function for it: it is inline in ConvAttention.forward, so
    oracle/loglik_torch.py restates it op by op), then the reference's OWN numba `b_mas`, unmodified, from oracle/_ref
    (falls back to the C/OpenMP port, and says so, only if that cannot be loaded), then attn_hard.sum(dim=1)."""

    def __init__(self):
        import torch
        from oracle import ref_mas
        self.ncores = os.cpu_count() or 1
        torch.set_num_threads(self.ncores)           # every host core, whatever OMP_NUM_THREADS says (torchrun sets it to 1)
        self.b_mas, self.why = ref_mas.load()
        self.kind = "reference" if self.b_mas is not None else "port"
        self.mas_threads = self.ncores
        if self.b_mas is not None:
            import numba
            try:
                numba.set_num_threads(min(self.ncores, numba.config.NUMBA_NUM_THREADS))
            except Exception:
                pass
            self.mas_threads = int(numba.get_num_threads())

    def mas(self, logits_np, tl_np, ml_np, threads=None):
        if self.b_mas is not None:
            if threads is not None:
                import numba
                numba.set_num_threads(threads)
            try:
                return self.b_mas(logits_np, tl_np, ml_np)            # mutates logits_np: callers pass a scratch copy
            finally:
                if threads is not None:
                    import numba
                    numba.set_num_threads(self.mas_threads)
        from oracle import mas as omas
        return omas.b_mas_with_durations(logits_np, tl_np, ml_np, threads or self.ncores)[0]

    def step(self, q, k, tl, ml, scale):
        from oracle import loglik_torch as olt
        soft, logits = olt.loglik(q, k, tl, ml, scale)
        hard = self.mas(logits.numpy(), tl.numpy(), ml.numpy())
        return hard.sum(axis=1, dtype=np.int64)                        # alignment.py:275

    def describe(self, w, n):
        what = ("the reference's numba b_mas (tts/modules/aligner/mas.py:30-35, unmodified, staged in oracle/_ref)" if self.kind == "reference"
                else f"C/OpenMP port of b_mas (oracle/mas_oracle.c) because: {self.why}")
        whole = "the whole batch" if n == w.batch else f"first {n} of {w.batch} utterances"
        return f"{whole} of {w.name}; fp32; torch-CPU log-likelihood ops (alignment.py:189-208 sequence, {self.ncores} threads) + {what} on {self.mas_threads} threads"


def cpu_inputs(w, n):
    import torch
    from isp_tts_b200 import synth
    tl, ml = synth.workload_lengths(w)
    tl, ml = tl[:n].copy(), ml[:n].copy()
    q, k = synth.encoded_pair(n, w.t1max, w.t2max, w.dim, tl, ml, w.seed + 1)
    return torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(tl), torch.from_numpy(ml), tl, ml


def cpu_baseline(w, reps, warm=1, budget_s=25.0):
    """Timed on the WHOLE batch of the workload (a cfg3 step is ~0.2 s of host time) unless one pass would blow the budget."""
    cpu = CpuPath()
    n = w.batch
    qt, kt, tlt, mlt, tl, ml = cpu_inputs(w, n)
    scale = w.dim ** -0.5
    t0 = time.perf_counter()
    cpu.step(qt, kt, tlt, mlt, scale)                                   # warm-up: numba JIT (~20 s the first time), torch threads
    first = time.perf_counter() - t0
    for _ in range(max(0, warm - 1)):
        cpu.step(qt, kt, tlt, mlt, scale)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu.step(qt, kt, tlt, mlt, scale)
        times.append(time.perf_counter() - t0)
        if sum(times) > budget_s:
            break
    best = min(times)
    # MAS alone, as SURVEY.md 8d asks: b_mas(x.copy(), ...) with the copy outside the timed region, all threads and one
    from oracle import loglik_torch as olt
    logits = olt.loglik(qt, kt, tlt, mlt, scale)[1].numpy()
    mas_all, mas_one = [], []
    for _ in range(3):
        x = logits.copy()
        t0 = time.perf_counter(); cpu.mas(x, tl, ml); mas_all.append(time.perf_counter() - t0)
    x = logits.copy()
    t0 = time.perf_counter(); cpu.mas(x, tl, ml, threads=1); mas_one.append(time.perf_counter() - t0)
    out = {
        "value": n / best, "unit": UNIT, "cores": int(cpu.ncores), "kind": cpu.kind, "sample": cpu.describe(w, n) + f"; best of {len(times)} (median {n / float(np.median(times)):.1f}); first call incl. JIT {first:.1f} s",
        "valid_cells_per_s": float((tl * ml).sum()) / best, "ms_per_step": best * 1e3,
        "mas_only": {"all_threads_utt_per_s": n / min(mas_all), "threads": cpu.mas_threads, "one_thread_utt_per_s": n / min(mas_one),
                     "ms_all_threads": min(mas_all) * 1e3},
    }
    return out, times, cpu, logits


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, times, cpu, _ = cpu_baseline(w, reps=max(1, args.steps), warm=max(1, min(args.warmup, 2)), budget_s=120.0)
    from isp_tts_b200 import synth
    tl, ml = synth.workload_lengths(w)
    med = float(np.median(times))
    cfg = base_config(w, tl, ml)
    cfg["step"] = "one pass of the hot path over the whole batch on the host cores (the reference's CPU route)"
    out = {
        "impl": "reference", "metric": METRIC, "value": w.batch / med, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": w.batch / med, "unit": UNIT, "cores": base["cores"], "kind": base["kind"], "sample": base["sample"],
                         "mas_only": base["mas_only"]},
        "e2e": {"value": w.batch / med, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def graph_ms(torch, fn, reps=10):
    """Device time of one call of `fn` (ms), measured over `reps` replays of a CUDA graph of it: the next-row kernels are tens of
    microseconds long while their Python wrappers (ctypes, three tensor-map encodes per GEMM) cost about as much on the host, so
    eager launches would time the host.  Falls back to eager timing if the capture fails."""
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        s.record()
        for _ in range(reps):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps
    except Exception as exc:                               # pragma: no cover
        print(f"[bench] graph capture of a next-row measurement failed ({exc!r}); eager timing", file=sys.stderr, flush=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    from isp_tts_b200 import synth
    from isp_tts_b200.alignment import _align_cuda, _loglik_cuda, pack_rows, stage_operands, unpack_operands
    from isp_tts_b200.mas import mas_forward

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.stage_ctas:
        from isp_tts_b200 import _lib as _l
        _l.set_option("stage.ctas", args.stage_ctas)
    for kv in args.opt:
        from isp_tts_b200 import _lib as _lo
        key, val = kv.split("=")
        _lo.set_option(key, int(val))
    if args.mas_ring or args.mas_slots:
        from isp_tts_b200 import _lib
        _lib.set_option("mas.ring_rows", args.mas_ring)
        _lib.set_option("mas.slots", args.mas_slots)
    gemm_dtype = torch.bfloat16 if args.gemm == "bf16" else torch.float32
    elem = 2 if args.gemm == "bf16" else 4
    B, T1, T2, D = w.batch, w.t1max, w.t2max, w.dim
    strong = w.name.startswith("cfg5")
    if strong:
        # ONE global batch, the same on every rank, dealt to the ranks by cell count (sharding.balanced_assignment): strong scaling
        from isp_tts_b200 import sharding
        tl_g, ml_g = synth.lengths(B, T2, T1, w.ragged, w.seed)
        mine = sharding.balanced_assignment(tl_g, ml_g, world)[rank]
        global_B, global_cells = B, int((tl_g * ml_g).sum())
        tl, ml = tl_g[mine].copy(), ml_g[mine].copy()
        B = len(mine)
        q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, w.seed + 1 + 1000 * rank)
    else:
        # every rank owns a different batch of the same law (utterance sharding, weak scaling)
        tl, ml = synth.lengths(B, T2, T1, w.ragged, w.seed + 1000 * rank)
        q, k = synth.encoded_pair(B, T1, T2, D, tl, ml, w.seed + 1 + 1000 * rank)
        global_B, global_cells = world * B, None
    scale = D ** -0.5
    q_host = torch.from_numpy(q).to(gemm_dtype).pin_memory()
    k_host = torch.from_numpy(k).to(gemm_dtype).pin_memory()
    tl_host, ml_host = torch.from_numpy(tl).pin_memory(), torch.from_numpy(ml).pin_memory()
    q_dev, k_dev = q_host.to(dev), k_host.to(dev)
    tl_dev, ml_dev = tl_host.to(dev), ml_host.to(dev)

    linked = args.link == "on"

    def step_resident(events=None):
        if linked and events is None:
            # the product's step: isp_align_forward (both kernels, the second starting under the first's last wave)
            return _align_cuda(q_dev, k_dev, tl_dev, ml_dev, scale, True)[3]
        if events is not None:
            events[0].record()
        soft, logits = _loglik_cuda(q_dev, k_dev, tl_dev, ml_dev, scale, True)
        if events is not None:
            events[1].record()
        hard, dur = mas_forward(logits, tl_dev, ml_dev)
        if events is not None:
            events[2].record()
        return dur

    # End to end through host buffers, double-buffered: the H2D copy of step i+1 runs on a copy stream under the kernels of
    # step i (every step still pays its own H2D of Q, K and the lengths and its own D2H of the durations inside the timed
    # region; PCIe, not the kernels, bounds this number).
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [dict(q=torch.empty_like(q_dev), k=torch.empty_like(k_dev), tl=torch.empty_like(tl_dev), ml=torch.empty_like(ml_dev),
                 ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]
    if args.e2e_copy == "packed":
        # the host holds the valid rows back to back (what a data loader that does not pad hands over): one plain DMA per tensor
        qp_host, kp_host = pack_rows(q_host, ml).pin_memory(), pack_rows(k_host, tl).pin_memory()
        for bf in bufs:
            bf["qp"], bf["kp"] = torch.empty_like(qp_host, device=dev), torch.empty_like(kp_host, device=dev)
    arena_host = None
    if args.e2e_copy == "arena":
        # the same packed rows and the two length vectors as views of ONE pinned arena (what a collate function that writes into a
        # pre-pinned buffer produces): one DMA per step, nothing between two steps' transfers on the copy stream; the scatter into
        # the padded operands runs on the compute stream
        qp_t, kp_t = pack_rows(q_host, ml), pack_rows(k_host, tl)
        parts = [("tl", tl_host), ("ml", ml_host), ("qp", qp_t), ("kp", kp_t)]
        offs, total_b = {}, 0
        for name, t in parts:
            offs[name] = total_b
            total_b += (t.numel() * t.element_size() + 255) // 256 * 256
        arena_host = torch.empty((total_b,), dtype=torch.uint8).pin_memory()

        def views(arena):
            return {name: arena[offs[name]: offs[name] + t.numel() * t.element_size()].view(t.dtype).view(t.shape) for name, t in parts}
        hv = views(arena_host)
        for name, t in parts:
            hv[name].copy_(t)
        for bf in bufs:
            bf["arena"] = torch.empty((total_b,), dtype=torch.uint8, device=dev)
            bf.update(views(bf["arena"]))
    dur_hosts = [torch.empty((B, T2), dtype=torch.int64).pin_memory() for _ in range(2)]
    hard_hosts = [torch.empty((B, T1, T2), dtype=torch.int16).pin_memory() for _ in range(2)] if args.e2e_outputs == "hard" else None
    state = {"i": 0}

    def e2e_upload(slot):
        bf = bufs[slot]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(bf["free"])                 # the kernels that read this buffer two steps ago are done
            if args.e2e_copy == "arena":
                bf["arena"].copy_(arena_host, non_blocking=True)
                bf["ready"].record(copy_stream)
                return
            bf["tl"].copy_(tl_host, non_blocking=True)
            bf["ml"].copy_(ml_host, non_blocking=True)
            if args.e2e_copy == "packed":
                # two copy-engine transfers of the packed rows, then a scatter into the padded operands on the device
                bf["qp"].copy_(qp_host, non_blocking=True)
                bf["kp"].copy_(kp_host, non_blocking=True)
                unpack_operands(bf["qp"], bf["kp"], bf["tl"], bf["ml"], T1, T2, out_q=bf["q"], out_k=bf["k"])
            elif args.e2e_copy == "staged":
                # ragged staging: only the rows below the lengths cross PCIe, the padding is zero-filled on the device
                stage_operands(q_host, k_host, bf["tl"], bf["ml"], out_q=bf["q"], out_k=bf["k"])
            else:
                bf["q"].copy_(q_host, non_blocking=True)
                bf["k"].copy_(k_host, non_blocking=True)
            bf["ready"].record(copy_stream)

    def step_e2e():
        i = state["i"]
        slot = i & 1
        if i == 0:
            e2e_upload(0)
        e2e_upload(slot ^ 1)                                   # next step's inputs, under this step's kernels
        bf = bufs[slot]
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(bf["ready"])
        if args.e2e_copy == "arena":
            unpack_operands(bf["qp"], bf["kp"], bf["tl"], bf["ml"], T1, T2, out_q=bf["q"], out_k=bf["k"])
        if linked:
            soft, logits, hard, dur, _ = _align_cuda(bf["q"], bf["k"], bf["tl"], bf["ml"], scale, True)
        else:
            soft, logits = _loglik_cuda(bf["q"], bf["k"], bf["tl"], bf["ml"], scale, True)
            hard, dur = mas_forward(logits, bf["tl"], bf["ml"])
        bf["free"].record(cur)
        dur_hosts[slot].copy_(dur, non_blocking=True)
        if hard_hosts is not None:
            hard_hosts[slot].copy_(hard, non_blocking=True)          # what the reference's CPU route hands back (alignment.py:312)
        state["i"] = i + 1
        return dur_hosts[slot]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- resident-input timing (value) --------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    # per-kernel times: an eager pass with events between the two launches (not the throughput measurement)
    n_ev = min(args.steps, 20)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n_ev)]
    for s in range(n_ev):
        step_resident(evs[s])
    barrier()
    t_loglik = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    t_mas = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    # The step is two kernels and a 256 B memset with fixed shapes and buffers: it is captured once into a CUDA graph and
    # replayed, so that the host's launch path (Python wrappers, ctypes, the caching allocator: ~0.1 ms per step, more
    # with eight ranks sharing the host cores) is not what the timed region measures.  --launch eager times the wrappers.
    graph = None
    launch = args.launch
    if launch == "graph":
        try:
            graph = torch.cuda.CUDAGraph()
            # captured on a stream the step has already run on: the linked call keeps one workspace per stream and shape, and a
            # workspace it has used before needs no memset in front of the kernels (ISP_ALIGN_WS_CLEAN)
            cap_stream = torch.cuda.Stream(device=dev)
            cap_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(cap_stream):
                step_resident()
            cap_stream.synchronize()
            with torch.cuda.graph(graph, stream=cap_stream):
                dur_static = step_resident()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            if not torch.equal(dur_static, step_resident()):
                raise RuntimeError("graph replay and eager launch disagree")
        except Exception as exc:                                    # report it, and time the eager launches instead
            print(f"[bench] CUDA graph capture failed ({exc!r}); timing eager launches", file=sys.stderr, flush=True)
            graph = None
            launch = "eager (graph capture failed)"
            torch.cuda.synchronize()
    # Per-kernel times the way the step is timed: the log-likelihood alone as a replayed graph, MAS as what it adds to the step
    # (the eager events above include the host's launch path, which at small batches is longer than the log-likelihood kernel
    # it should hide under: cfg2's MAS read 57 us there against 41 us launched back to back)
    t_eager = (t_loglik, t_mas)
    kernel_timing = "CUDA events around eager launches"
    if graph is not None:
        try:
            t_ll_g = graph_ms(torch, lambda: _loglik_cuda(q_dev, k_dev, tl_dev, ml_dev, scale, True), reps=max(10, min(args.steps, 50)))
            st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st.record()
            for _ in range(max(10, min(args.steps, 50))):
                graph.replay()
            en.record()
            torch.cuda.synchronize()
            t_step_g = st.elapsed_time(en) / max(10, min(args.steps, 50))
            if 0.0 < t_ll_g < t_step_g:
                t_loglik, t_mas = t_ll_g, t_step_g - t_ll_g
                kernel_timing = ("graph replay: isp_loglik alone; isp_mas = step - isp_loglik (what it adds to the step)"
                                 + ("; the step is isp_align_forward: the MAS kernel starts under the log-likelihood kernel's last wave, "
                                    "launched on its own (isp_mas_forward) it takes ms_eager_events.isp_mas" if linked else ""))
        except Exception as exc:                                    # pragma: no cover
            print(f"[bench] per-kernel graph timing failed ({exc!r}); keeping the eager events", file=sys.stderr, flush=True)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        start.record()
        if graph is not None:
            for s in range(args.steps):
                graph.replay()
        else:
            for s in range(args.steps):
                step_resident()
        end.record()
        barrier()
    total_ms = max_over_ranks(start.elapsed_time(end))

    # ---- end-to-end timing through host buffers (e2e) -----------------------------------
    for _ in range(3):
        step_e2e()
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk2:
        barrier()
        s2.record()
        for _ in range(args.steps):
            step_e2e()
        e2.record()
        barrier()
    e2e_ms = max_over_ranks(s2.elapsed_time(e2))
    # The platform's ceiling for that leg: the same number of bytes per step as ONE plain pinned cudaMemcpyAsync on the copy
    # engine, all ranks at once, nothing else running (what the host side of the PCIe tree gives N GPUs together)
    ceil_bytes = int((ml.sum() + tl.sum()) * D * q_host.element_size()) if args.e2e_copy in ("staged", "packed", "arena") else \
        q_host.numel() * q_host.element_size() + k_host.numel() * k_host.element_size()
    flat_host = q_host.view(-1)[: ceil_bytes // q_host.element_size()]
    flat_dev = torch.empty_like(flat_host, device=dev)
    for _ in range(2):
        flat_dev.copy_(flat_host, non_blocking=True)
    barrier()
    s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s3.record()
    for _ in range(10):
        flat_dev.copy_(flat_host, non_blocking=True)
    e3.record()
    barrier()
    ceil_ms = max_over_ranks(s3.elapsed_time(e3)) / 10
    link_ceiling_gbs = flat_host.numel() * flat_host.element_size() * world / ceil_ms / 1e6
    del flat_dev
    # durations must be what the resident path produced
    ref_dur = step_resident().cpu()
    torch.cuda.synchronize()
    dur_host = dur_hosts[(state["i"] - 1) & 1]
    if not torch.equal(ref_dur, dur_host):
        raise RuntimeError("e2e durations differ from the resident-input run")
    if int(dur_host.sum()) != int(ml.sum()):
        raise RuntimeError("durations do not sum to the mel lengths")

    # ---- the one collective of the path (outside every timed region): NCCL all-gather of the durations when the caller
    # wants a single tensor; checked against the lengths every rank holds ---------------------------
    gather = None
    if world > 1:
        from isp_tts_b200 import sharding
        dur_dev = step_resident()
        gs_, ge_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        counts = [len(a) for a in sharding.balanced_assignment(tl_g, ml_g, world)] if strong else [B] * world
        off = sum(counts[:rank])
        sharding.gather_durations(dur_dev, t2max=T2, counts=counts)           # warm-up (NCCL communicator, buffers)
        gs_.record()
        full = sharding.gather_durations(dur_dev, t2max=T2, counts=counts)
        ge_.record()
        torch.cuda.synchronize()
        frames = torch.tensor([int(ml.sum())], dtype=torch.int64, device=dev)
        dist.all_reduce(frames)
        if tuple(full.shape) != (sum(counts), T2) or int(full.sum()) != int(frames.item()) or not torch.equal(full[off:off + B], dur_dev):
            raise RuntimeError("gathered durations do not match the ranks' own")
        if strong and int(full.sum()) != int(ml_g.sum()):
            raise RuntimeError("gathered durations of the global batch do not sum to its mel lengths")
        gather = {"collective": "ncclAllGather of durations (B_local, T2max) int64 via torch.distributed", "ms": gs_.elapsed_time(ge_),
                  "bytes_per_rank": int(dur_dev.numel() * 8), "shape": list(full.shape)}

    # ---- next row: backward of the log-likelihood, measured beside the hot path ----------
    bwd = None
    if rank == 0 and not args.no_backward:
        from isp_tts_b200.alignment import _scores, loglik_backward_ds, loglik_backward_from_logits
        from isp_tts_b200.gemm import bgemm
        soft_b, logits_b, rowsum_b = _loglik_cuda(q_dev, k_dev, tl_dev, ml_dev, scale, True, want_rowsum=True)
        gen = torch.Generator(device=dev).manual_seed(1)
        g_l = torch.randn(soft_b.shape, device=dev, generator=gen)
        g_s = torch.randn(soft_b.shape, device=dev, generator=gen)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        box = {}

        def f_scores():
            box["sc"] = _scores(q_dev, k_dev, tl_dev, ml_dev)                            # isp_gemm_batched, both operands K-major

        def f_ds():
            box["ds"] = loglik_backward_ds(box["sc"], soft_b, g_l, g_s, scale, True, out_dtype=gemm_dtype)

        def f_ds2():
            # the product's route: dS from attn_logits and the prior's saved row sums -- no score GEMM, no attn_soft read
            box["ds2"] = loglik_backward_from_logits(logits_b, g_l, g_s, rowsum_b, tl_dev, ml_dev, scale, True, out_dtype=gemm_dtype)

        def f_grads():
            box["gq"] = bgemm(box["ds"], k_dev, out_dtype=gemm_dtype, k_len=tl_dev)                  # dQ = dS.K   (K read MN-major)
            box["gk"] = bgemm(box["ds"].transpose(1, 2), q_dev, out_dtype=gemm_dtype, k_len=ml_dev)  # dK = dS^T.Q (both MN-major)

        def f_lib():
            box["sc_l"] = torch.bmm(q_dev, k_dev.transpose(1, 2), out_dtype=torch.float32) if gemm_dtype != torch.float32 else torch.matmul(q_dev, k_dev.transpose(1, 2))
            box["gq_l"] = torch.matmul(box["ds"], k_dev)
            box["gk_l"] = torch.matmul(box["ds"].transpose(1, 2), q_dev)

        def f_all_scores():
            f_scores(); f_ds(); f_grads()

        from_logits = rowsum_b is not None and T2 % 4 == 0

        def f_all():
            if from_logits:
                f_ds2(); box["ds"] = box["ds2"]; f_grads()
            else:
                f_all_scores()

        t_s, t_ds, t_mm, t_lib = ([graph_ms(torch, f)] for f in (f_scores, f_ds, f_grads, f_lib))
        t_all_scores = graph_ms(torch, f_all_scores)
        t_ds2 = graph_ms(torch, f_ds2) if from_logits else None
        if from_logits:
            dd = float((box["ds2"].float() - box["ds"].float()).abs().max() / box["ds"].float().abs().max())
            if dd > 2e-2:
                raise RuntimeError("dS from attn_logits differs from dS from the recomputed scores: %g" % dd)
        t_all = graph_ms(torch, f_all)
        sc, d_s, gq, gk, gq_l, gk_l = box["sc"], box["ds"], box["gq"], box["gk"], box["gq_l"], box["gk_l"]
        gerr = max(float((gq.float() - gq_l.float()).abs().max() / gq_l.float().abs().max()),
                   float((gk.float() - gk_l.float()).abs().max() / gk_l.float().abs().max()))
        if gerr > 2e-2:
            raise RuntimeError("dQ / dK of isp_gemm_batched differ from the library GEMMs: %g" % gerr)
        by_ds = (16 + elem) * B * T1 * T2
        fl_bwd = 3 * 2 * D * B * T1 * T2
        bwd = {"row": "f-1 backward of the log-likelihood (d attn_logits, d attn_soft -> dQ, dK): no library GEMM",
               "isp_loglik_backward_ds": {"ms": float(np.mean(t_ds)), "algorithmic_bytes": by_ds,
                                          "gbs": by_ds / float(np.mean(t_ds)) / 1e6},
               "isp_gemm_batched_ms": {"scores_QKt": float(np.mean(t_s)), "dQ_and_dK": float(np.mean(t_mm))},
               "library_gemms_same_three_products_ms": float(np.mean(t_lib)),
               "max_rel_diff_vs_library": gerr, "padded_flops": fl_bwd,
               "total_ms": float(t_all), "timing": "CUDA graph replay of each part and of the whole backward",
               "route": ("isp_loglik_backward_from_logits + dQ, dK (what _LogLikelihood.backward runs)" if from_logits
                         else "scores + isp_loglik_backward_ds + dQ, dK"),
               "total_ms_scores_route": float(t_all_scores)}
        if from_logits:
            by2 = (12 + elem) * B * T1 * T2
            bwd["isp_loglik_backward_from_logits"] = {"ms": float(t_ds2), "algorithmic_bytes": by2, "gbs": by2 / float(t_ds2) / 1e6,
                                                      "max_rel_diff_vs_scores_route": dd}
        box.clear()
        del gq_l, gk_l
        # f-3: binarization loss from the path vs the reference's boolean-mask gather on the dense tensors (loss.py:97-105)
        from isp_tts_b200.mas import binarization_loss
        hard_b, dur_b, path_b = mas_forward(logits_b, tl_dev, ml_dev, return_path=True)
        t_ours, t_ref = [], []
        for it in range(6):
            ev[0].record()
            l1 = binarization_loss(soft_b, path_b, ml_dev)
            ev[1].record()
            l2 = -torch.log(torch.clamp(soft_b[hard_b == 1], min=1e-6)).sum() / hard_b.sum()
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                t_ours.append(ev[0].elapsed_time(ev[1])); t_ref.append(ev[1].elapsed_time(ev[2]))
        if abs(float(l1) - float(l2)) > 1e-4 * max(1.0, abs(float(l2))):
            raise RuntimeError("binarization loss from the path differs from the dense formula")
        bwd["f-3 binarization loss"] = {"isp_bin_loss_sums_ms": float(np.mean(t_ours)), "torch_dense_mask_ms": float(np.mean(t_ref)),
                                        "value": float(l1)}
        # f-3: MAS for consumers that take the path: no dense attn_hard (isp_mas_forward_path with attn_hard == NULL)
        t_d, t_p = [], []
        for it in range(8):
            ev[0].record()
            mas_forward(logits_b, tl_dev, ml_dev, return_path=True)
            ev[1].record()
            mas_forward(logits_b, tl_dev, ml_dev, return_path=True, dense=False)
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 3:
                t_d.append(ev[0].elapsed_time(ev[1])); t_p.append(ev[1].elapsed_time(ev[2]))
        bwd["f-3 MAS without the dense output"] = {"isp_mas_forward_path_ms": float(np.mean(t_d)), "attn_hard_null_ms": float(np.mean(t_p)),
                                                   "bytes_not_written": 2 * B * T1 * T2}
        # f-3: length regulator from the path vs the reference's matmul with a (T1 x T2) 0/1 matrix (temporal_adaptor.py:420-431)
        from isp_tts_b200.consumers import length_regulate
        enc_dim = 384                                   # recipes/acoustic/core.yaml: encoder dim
        xe = torch.randn((B, T2, enc_dim), device=dev, generator=gen)
        t_ours, t_ref = [], []
        for it in range(6):
            ev[0].record()
            o1 = length_regulate(xe, path_b, dur_b)
            ev[1].record()
            reps = (dur_b.float() + 0.5).long()
            cums = torch.cumsum(torch.nn.functional.pad(reps, (1, 0, 0, 0), value=0.0), dim=1, dtype=xe.dtype)[:, None, :]
            r = torch.arange(int(T1), device=dev)[None, :, None]
            o2 = torch.matmul(((cums[:, :, :-1] <= r) & (cums[:, :, 1:] > r)).to(xe.dtype), xe)
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                t_ours.append(ev[0].elapsed_time(ev[1])); t_ref.append(ev[1].elapsed_time(ev[2]))
        if not torch.equal(o1, o2):
            raise RuntimeError("length regulator from the path differs from the reference formula")
        by_lr = 4 * B * T1 * enc_dim + 2 * B * T1
        bwd["f-3 length regulator"] = {"isp_length_regulate_ms": float(np.mean(t_ours)), "algorithmic_bytes": by_lr,
                                       "gbs": by_lr / float(np.mean(t_ours)) / 1e6, "torch_reference_ms": float(np.mean(t_ref)),
                                       "shape": f"x ({B}, {T2}, {enc_dim}) fp32 -> ({B}, {T1}, {enc_dim})"}
        # f-3, soft route (the recipe's default, core.yaml:148): LengthRegulator as alignment @ x (temporal_adaptor.py:417-419) on the
        # tcgen05 batched GEMM, and TemporalAverager as one stream over the alignment (:446-449)
        from isp_tts_b200.soft import soft_average, soft_expand
        def f_se():
            box["o3"] = soft_expand(soft_b, xe, frame_len=ml_dev, token_len=tl_dev)

        def f_se_ref():
            box["o4"] = (xe.transpose(1, 2) @ soft_b.transpose(1, 2)).transpose(1, 2)

        t_ours, t_ref = [graph_ms(torch, f_se)], [graph_ms(torch, f_se_ref)]
        o3, o4 = box["o3"], box["o4"]
        serr = float((o3 - o4).abs().max() / o4.abs().max())
        if serr > 5e-3:
            raise RuntimeError("soft_expand differs from the reference formula: %g" % serr)
        by_se = 4 * int((ml * tl).sum()) + 4 * B * T2 * enc_dim + 4 * B * T1 * enc_dim
        bwd["f-3 soft length regulator"] = {"isp_gemm_batched_ms": float(np.mean(t_ours)), "torch_reference_ms": float(np.mean(t_ref)),
                                            "algorithmic_bytes": by_se, "gbs": by_se / float(np.mean(t_ours)) / 1e6,
                                            "padded_flops": 2 * B * T1 * T2 * enc_dim, "max_rel_diff": serr,
                                            "shape": f"attn_soft ({B}, {T1}, {T2}) fp32 (TF32 products) @ x ({B}, {T2}, {enc_dim})"}
        feat = torch.rand((B, 2, T1), device=dev, generator=gen) * 200.0
        def f_sa():
            box["a1"] = soft_average(feat, soft_b, row_len=ml_dev)

        def f_sa_ref():
            box["a2"] = feat @ soft_b / (soft_b.sum(dim=1, keepdim=True) + 1e-5)

        t_ours, t_ref = [graph_ms(torch, f_sa)], [graph_ms(torch, f_sa_ref)]
        a1, a2 = box["a1"], box["a2"]
        if not torch.allclose(a1, a2, rtol=1e-3, atol=1e-2):
            raise RuntimeError("soft_average differs from the reference formula")
        by_sa = 4 * int(ml.sum()) * T2
        bwd["f-3 soft averager"] = {"isp_soft_average_ms": float(np.mean(t_ours)), "torch_reference_ms": float(np.mean(t_ref)),
                                    "algorithmic_bytes": by_sa, "gbs": by_sa / float(np.mean(t_ours)) / 1e6}
        del g_l, g_s, sc, d_s, gq, gk, hard_b, dur_b, path_b, xe, o1, o2, o3, o4, a1, a2, feat, soft_b
        # f-2: the projection stacks at the recipe's widths (mel 80, text 384, attention_dim 128, kernels 5; 1.7 M parameters) on
        # this workload's batch shape: fused sm_100a route (stacks.py) against the torch ops under the same autocast
        from isp_tts_b200 import Aligner
        al = Aligner(**synth.RECIPE_HP).eval().to(dev)
        al.attention.gemm_dtype = "bf16"
        mel_np, txt_np = synth.recipe_inputs(w.seed + 7, B, T1, T2, tl, ml)
        mel_d, txt_d = torch.from_numpy(mel_np).to(dev), torch.from_numpy(txt_np).to(dev)
        def f_st():
            al.attention.fused_stacks = True
            with torch.no_grad():
                box["qf"], box["kf"] = al.attention.encode(mel_d, txt_d, ml_dev, tl_dev)

        def f_st_ref():
            al.attention.fused_stacks = False
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                box["qt"], box["kt"] = al.attention.encode(mel_d, txt_d, ml_dev, tl_dev)

        t_ours, t_ref = [graph_ms(torch, f_st, reps=5)], [graph_ms(torch, f_st_ref, reps=5)]
        qf, kf, qt, kt = box["qf"], box["kf"], box["qt"], box["kt"]
        box.clear()
        qerr = float((qf.float() - qt.float()).abs().max() / qt.float().abs().max())
        fl_pad = 2 * B * (T2 * (768 * 384 * 5 + 128 * 768) + T1 * (160 * 80 * 5 + 80 * 160 * 5 + 128 * 80))
        rows_k, rows_q = int(((tl + 127) // 128 * 128).sum()), int(((ml + 127) // 128 * 128).sum())
        fl_exec = 2 * (rows_k * (768 * 384 * 5 + 128 * 768) + rows_q * (160 * 80 * 5 + 80 * 160 * 5 + 128 * 80))
        bwd["f-2 projection stacks"] = {"fused_sm100_ms": float(np.mean(t_ours)), "torch_ops_autocast_ms": float(np.mean(t_ref)),
                                        "padded_flops": fl_pad, "executed_flops_valid_tiles": fl_exec,
                                        "tflops_executed": fl_exec / float(np.mean(t_ours)) / 1e9, "q_max_rel_diff_vs_torch": qerr,
                                        "kernels": "isp_prep_channels_last, isp_gemm_batched (implicit-GEMM conv, GELU + norm sums fused), isp_instance_norm_apply",
                                        "inner_dtype": "float16 activations and weights, bf16 Q / K out"}
        del al, mel_d, txt_d, qf, kf, qt, kt
        # f-4: forward-sum (CTC) loss and its gradient vs the reference's op sequence on the GPU (loss.py:59-79:
        # F.pad + log_softmax + transpose + nn.CTCLoss(zero_infinity=True)), same logits
        from isp_tts_b200.ctc import attention_ctc_loss
        from isp_tts_b200._lib import IspError
        t_ours, t_ref = [], []
        tgt = torch.arange(1, T2 + 1, device=dev)[None].expand(B, -1).clone()
        tgt[tgt > tl_dev[:, None]] = 0
        for it in range(5):
            xa = logits_b.detach().clone().requires_grad_(True)
            xb2 = logits_b.detach().clone().requires_grad_(True)
            ev[0].record()
            try:
                la = attention_ctc_loss(xa, tl_dev, ml_dev)
            except IspError as exc:                      # a shape the next-row kernel does not cover: say so, measure the rest
                bwd["f-4 forward-sum (CTC) loss"] = {"unsupported": str(exc)}
                break
            la.backward()
            ev[1].record()
            lp = torch.nn.functional.log_softmax(torch.nn.functional.pad(xb2, (1, 0), value=-1.0), dim=2).transpose(0, 1)
            lb2 = torch.nn.functional.ctc_loss(lp, tgt, ml_dev, tl_dev, blank=0, reduction="mean", zero_infinity=True)
            lb2.backward()
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                t_ours.append(ev[0].elapsed_time(ev[1])); t_ref.append(ev[1].elapsed_time(ev[2]))
        if "f-4 forward-sum (CTC) loss" not in bwd:
            if abs(la.item() - lb2.item()) > 1e-4 * max(1.0, abs(lb2.item())):
                raise RuntimeError("forward-sum loss differs from torch CTC: %r vs %r" % (la.item(), lb2.item()))
            gerr = float((xa.grad - xb2.grad).abs().max() / xb2.grad.abs().max())
            bwd["f-4 forward-sum (CTC) loss"] = {"isp_ctc_forward_backward_ms": float(np.mean(t_ours)),
                                                 "torch_reference_sequence_ms": float(np.mean(t_ref)),
                                                 "value": la.item(), "grad_max_rel_diff_vs_torch_fp32": gerr}
        del logits_b, xa, xb2

        # ---- the same step at the reference's own precision outside autocast: fp32 operands, true-fp32 products ----
        try:
            from isp_tts_b200.alignment import _align_cuda as _al
            q32, k32 = q_dev.float(), k_dev.float()
            t_tf32 = graph_ms(torch, lambda: _al(q32, k32, tl_dev, ml_dev, scale, True, precision="tf32"))
            t_fp32 = graph_ms(torch, lambda: _al(q32, k32, tl_dev, ml_dev, scale, True, precision="fp32"))
            l_t = _al(q32, k32, tl_dev, ml_dev, scale, True, precision="tf32")[1]
            l_f = _al(q32, k32, tl_dev, ml_dev, scale, True, precision="fp32")[1]
            bwd["precision rows (same workload, fp32 operands)"] = {
                "fp32_faithful_3xtf32_ms": t_fp32, "fp32_faithful_utt_per_s": B / t_fp32 * 1e3,
                "tf32_fused_ms": t_tf32, "tf32_fused_utt_per_s": B / t_tf32 * 1e3,
                "max_abs_diff_logits_tf32_vs_fp32": float((l_t - l_f).abs().max()),
                "what": "isp_split_3xtf32 x2 + isp_gemm_batched (TF32 over 3 D) + isp_loglik_rows + isp_mas_forward  vs  isp_align_forward on fp32 "
                        "operands (one TF32 product per term); the headline above is BASELINE configs[2]'s bf16 GEMM"}
            del q32, k32, l_t, l_f
        except Exception as exc:                                # pragma: no cover
            bwd["precision rows (same workload, fp32 operands)"] = {"unsupported": repr(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, tf_peak, which = load_peaks()
    if bwd is not None:
        bwd["isp_loglik_backward_ds"]["hbm_frac"] = bwd["isp_loglik_backward_ds"]["gbs"] / hbm_peak
        if "isp_loglik_backward_from_logits" in bwd:
            bwd["isp_loglik_backward_from_logits"]["hbm_frac"] = bwd["isp_loglik_backward_from_logits"]["gbs"] / hbm_peak
        bwd["f-3 length regulator"]["hbm_frac"] = bwd["f-3 length regulator"]["gbs"] / hbm_peak
        bwd["f-3 soft length regulator"]["hbm_frac"] = bwd["f-3 soft length regulator"]["gbs"] / hbm_peak
        bwd["f-3 soft averager"]["hbm_frac"] = bwd["f-3 soft averager"]["gbs"] / hbm_peak
        bwd["f-2 projection stacks"]["tensor_frac_executed"] = bwd["f-2 projection stacks"]["tflops_executed"] / tf_peak
    by_mas = mas_bytes(tl, ml, B, T1, T2)
    by_ll = loglik_bytes(B, T1, T2, D, elem)
    fl_ll = loglik_flops(B, T1, T2, D)
    # the log-likelihood kernel only reads operand rows below the lengths, and only runs the MMA on 128-frame tiles that hold a
    # valid frame: both accountings are reported (SURVEY.md 8d's padded-operand formula, and what really has to move / execute)
    by_ll_valid = elem * D * int(ml.sum() + tl.sum()) + 8 * B * T1 * T2
    fl_ll_exec = 2 * D * int(((ml + 127) // 128 * 128).sum()) * ((T2 + 15) // 16 * 16)
    ncu = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")       # per-launch DRAM bytes / pipe activity from the committed ncu capture
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                ncu = json.load(f).get(w.name.split(":")[0], {})
        except Exception:
            ncu = {}
    kern = {
        "isp_loglik (tcgen05 GEMM + fused epilogue)": {
            "ms": t_loglik, "algorithmic_bytes": by_ll, "gbs": by_ll / t_loglik / 1e6, "hbm_frac": by_ll / t_loglik / 1e6 / hbm_peak,
            "valid_row_bytes": by_ll_valid, "hbm_frac_valid_rows": by_ll_valid / t_loglik / 1e6 / hbm_peak,
            "tflops_padded": fl_ll / t_loglik / 1e9, "tflops_executed": fl_ll_exec / t_loglik / 1e9,
            "tensor_frac_executed": fl_ll_exec / t_loglik / 1e9 / tf_peak,
            "tensor_pipe_active_ncu": ncu.get("isp_loglik_tensor_pipe_active")},
        "isp_mas (wavefront DP + backtrack + durations)": {
            "ms": t_mas, "algorithmic_bytes": by_mas, "gbs": by_mas / t_mas / 1e6, "hbm_frac": by_mas / t_mas / 1e6 / hbm_peak},
    }
    kern["timing"] = kernel_timing
    kern["ms_eager_events"] = {"isp_loglik": t_eager[0], "isp_mas": t_eager[1]}
    dom = max((n for n in kern if isinstance(kern[n], dict) and "hbm_frac" in kern[n]), key=lambda n: kern[n]["ms"])

    def roof(name):
        return {"kernel": name, "bound": "hbm", "achieved": kern[name]["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": kern[name]["hbm_frac"],
                "traffic": ncu.get(name.split(" ")[0]), "peak_source": which + " (MEASURED_PEAKS.json hbm_gbs)"}
    roofline = roof(dom)
    roofline_all = [roof(n) for n in kern if isinstance(kern[n], dict) and "hbm_frac" in kern[n]]

    cpu = None
    if not args.no_cpu and world == 1:
        cpu, _, cpu_path, logits_np = cpu_baseline(w, reps=3)
        # the reference's end-to-end CPU route for MAS on a CUDA tensor (alignment.py:305-312): D2H fp32 + b_mas + H2D int16
        lg = torch.from_numpy(logits_np).to(dev)
        tl_np, ml_np = tl.copy(), ml.copy()
        route = []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = torch.from_numpy(cpu_path.mas(lg.detach().cpu().numpy(), tl_np, ml_np)).to(dev)
            torch.cuda.synchronize()
            route.append(time.perf_counter() - t0)
        cpu["reference_cpu_route_mas"] = {"ms": min(route) * 1e3, "what": "attn_logits.cpu().numpy() -> b_mas -> torch.from_numpy(...).to(device) (alignment.py:305-312)"}
        del lg, out

    utts = global_B * args.steps
    valid_cells = float(global_cells if strong else (tl * ml).sum() * world) * args.steps
    if args.e2e_copy in ("staged", "packed", "arena"):
        h2d = int((ml.sum() + tl.sum()) * D * q_host.element_size()) + 16 * B
    else:
        h2d = q_host.numel() * q_host.element_size() + k_host.numel() * k_host.element_size() + 16 * B
    out = {
        "metric": METRIC, "value": utts / (total_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "bf16 GEMM operands; f32 accumulate, epilogue and MAS" if elem == 2 else "tf32 GEMM products; f32 elsewhere",
        "data": "synthetic",
        "config": {**base_config(w, tl, ml),
                   "sharding": (f"one global batch of {global_B} utterances dealt to {world} rank(s) by cell count (longest first), "
                                f"{B} on rank 0; no collective on the data path") if strong else
                               f"by utterance, {world} rank(s), no collective on the data path",
                   "launch": launch if launch != "graph" else "CUDA graph replay of the step (isp_loglik_forward + isp_mas_forward)",
                   "l2": "per-step working set (operands + 3 dense outputs) = %.0f MB > 126 MB L2; no explicit flush" % (
                       (by_ll + 2 * B * T1 * T2) / 1e6)},
        "valid_cells_per_s": valid_cells / (total_ms / 1e3),
        "padded_cells_per_s": global_B * T1 * T2 * args.steps / (total_ms / 1e3),
        "roofline": roofline, "roofline_kernels": roofline_all, "kernels": kern, "cpu_baseline": cpu,
        "e2e": {"value": utts / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(dur_host.numel() * 8) + (2 * B * T1 * T2 if hard_hosts is not None else 0), "ms_per_step": e2e_ms / args.steps,
                "h2d_gbs_all_ranks": h2d * world / (e2e_ms / args.steps) / 1e6,
                "link_ceiling_gbs": link_ceiling_gbs,
                "link_ceiling": "aggregate H2D rate of plain pinned cudaMemcpyAsync of the same bytes on all ranks at once, nothing else running",
                "api": ({"staged": "isp_stage_operands (padded pinned host tensors; valid rows only cross PCIe, zero-copy reads) + ",
                         "packed": "two cudaMemcpyAsync of PACKED pinned host buffers (valid rows back to back) + isp_unpack_operands + ",
                         "arena": "ONE cudaMemcpyAsync per step of a pinned host arena holding the lengths and the PACKED rows of Q and K (valid rows back to "
                                  "back, as a collate function that writes into a pre-pinned buffer leaves them) + isp_unpack_operands on the compute stream + ",
                         "padded": ""}[args.e2e_copy])
                       + ("isp_align_forward (the two kernels linked)" if linked else "isp_loglik_forward + isp_mas_forward") + " through isp_tts_b200.  In: pinned host Q, K (already cast to the GEMM's "
                       + ("bf16" if elem == 2 else "fp32") + " on the host, outside the timed region) and int64 lengths.  Out: the int64 durations"
                       + (" and the dense int16 attn_hard (the reference's CPU route hands back attn_hard, alignment.py:312)" if hard_hosts is not None
                          else " only; attn_hard, attn_logits and attn_soft stay on the device (--e2e-outputs hard also brings attn_hard back)")
                       + ".  The next step's H2D overlaps this step's kernels (two device buffers, one copy stream)"},
        "step_api": "isp_align_forward" if linked else "isp_loglik_forward + isp_mas_forward",
        "next_rows": bwd, "durations_gather": gather,
        "gpu_launches": (2 + (1 if B > 512 else 0)) * args.steps,      # loglik + MAS (+ the MAS plan kernel beyond 512 utterances)
        "clocks": clk.summary(), "clocks_e2e": clk2.summary(),
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--gemm", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--mas-ring", type=int, default=0, help="tuning: MAS logit rows in flight, 0 = heuristic")
    ap.add_argument("--mas-slots", type=int, default=0, help="tuning: utterances per CTA (1|2), 0 = heuristic")
    ap.add_argument("--link", default="on", choices=["on", "off"],
                    help="on: the step is isp_align_forward (log-likelihood and MAS kernels linked by per-utterance ready counts); "
                         "off: isp_loglik_forward then isp_mas_forward")
    ap.add_argument("--opt", action="append", default=[], help="tuning: isp_set_option key=value (repeatable), e.g. mas.pdl=0")
    ap.add_argument("--stage-ctas", type=int, default=0, help="tuning: CTAs of the host->device staging kernel, 0 = default")
    ap.add_argument("--no-backward", action="store_true", help="skip the f-1 backward measurement that follows the timed steps")
    ap.add_argument("--batch", type=int, default=0, help="utterances per GPU (cfg5 sweep: 64..4096 with the cfg3 length law); 0 = the workload's own")
    ap.add_argument("--launch", default="graph", choices=["graph", "eager"],
                    help="resident-input timing: replay the step as a CUDA graph (default) or launch it through the Python wrappers")
    ap.add_argument("--e2e-copy", default="arena", choices=["arena", "packed", "staged", "padded"],
                    help="end-to-end H2D of Q and K: packed rows by DMA + isp_unpack_operands (default), isp_stage_operands (zero-copy reads "
                         "of the valid rows of padded host tensors), or plain copies of the padded tensors")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (sweeps)")
    ap.add_argument("--e2e-outputs", default="durations", choices=["durations", "hard"],
                    help="what the end-to-end leg reads back: the durations (default) or the durations and the dense attn_hard")
    args = ap.parse_args()
    from isp_tts_b200 import synth
    w = synth.WORKLOADS[args.workload]
    if args.batch > 0:
        import dataclasses
        w = dataclasses.replace(w, batch=args.batch, name=w.name.replace("batch 256", f"batch {args.batch}") + f" [batch {args.batch}]")
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
