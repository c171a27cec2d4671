#!/usr/bin/env python
"""bench.py — the hot path's headline metric on B200 (contract in the task statement).

A "step" is one engine tick + one render of the north-star workload, BASELINE.json configs[3]:
2^24 particle-style entities per GPU (Transform + Velocity + Sprite, 4 textures, 1 layer):
velocity system -> Transform->Matrix4 compose -> stable sort on (layer, texture, z) -> instance +
quad-vertex emission + batch table. Every rank owns a contiguous entity-id range of the same
seeded scene (weak scaling: 2^24 entities per rank, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

value  = entities/s, all ranks, inputs resident in HBM, K steps between CUDA events (max over ranks)
e2e    = the same through the C ABI with host buffers: H2D of every component array and D2H of
         instances, vertices, order and batch table inside the timed region, every step
roofline = the emit kernel (dominant): algorithmic bytes / its CUDA-event time vs measured HBM peak
cpu_baseline = the CPU oracle (restated reference path, 1 thread) on a bounded sample
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "entities/sec per frame (transform+batch)"
UNIT = "entities/s"
N_PER_GPU = 1 << 24
ALG_BYTES_FRAME = 232   # SURVEY.md §8d, particles: 48+16+12 read, 80+64+12 written
ALG_BYTES_EMIT = 208    # the emit kernel's share: Transform+Sprite read, Instance+4 QuadVertex written
CPU_SAMPLE = 1 << 22


def bench_config(n, world):
    """The `config` object of the bench line: the same in both arms (this build and --impl reference)."""
    return {"workload": workload_name(n), "entities_per_gpu": n, "sharding": f"contiguous id range x{world}",
            "l2": "no flush needed: each pass streams >= 0.5 GB per GPU, far larger than the 126 MB L2"}


def workload_name(n):
    return (f"configs[3]: {n} particle-style entities per GPU (Transform+Velocity+Sprite, 4 textures, 1 layer, "
            "dt=0.01): velocity system + transform + sprite batching per frame")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
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

    def _run(self):
        nv = self.nv
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksThrottleReason") or k.startswith("nvmlClocksEventReason"):
                v = getattr(nv, k)
                if isinstance(v, int) and v:
                    names.setdefault(v, k.replace("nvmlClocksThrottleReason", "").replace("nvmlClocksEventReason", ""))
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit and bit & (bit - 1) == 0:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        reasons = sorted(r for r in self.reasons if r not in ("GpuIdle", "None", "ApplicationsClocksSetting"))
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Runs this process on the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned buffer is allocated (first
    touch then places the pages there): with several ranks per box the host copies of the e2e leg otherwise all cross
    the socket interconnect from node 0. Returns what was done, for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return {"numa_node": None, "note": "the platform reports no NUMA node for this GPU"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"numa_node": node, "note": "none of the node's CPUs is available to this process"}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "note": f"not bound: {type(e).__name__}"}


def cpu_reference(steps, warmup, n_sample):
    """The CPU oracle (oracle/: restated tuber-ecs query + tuber-math compose + SPEC batching), single
    thread like the reference, on the first n_sample entities of the same scene."""
    from oracle import oracle
    from tuber_legacy_b200 import scenes
    sc = scenes.config4(n=N_PER_GPU, count=n_sample)
    ref = oracle.Scene(sc.n, sc.transforms, sc.sprites, sc.velocities)
    times = []
    for i in range(warmup + steps):
        t = ref.step(sc.dt)
        t += ref.frame(bounds=False).seconds
        if i >= warmup:
            times.append(t)
    ref.close()
    total = sum(times)
    return {"value": n_sample * len(times) / total, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"first {n_sample} entities of the workload scene, {len(times)} frames "
                      f"(velocity system + render_scene), {total:.1f} s of CPU time"}, total / len(times)


def config_table(capi, scenes, torch, device, peak, steps, which=None):
    """Every BASELINE config on ONE GPU x what happens between frames, device-resident, CUDA events around `steps`
    frames after 4 warm-up frames (every pass streams far more than the L2 holds, except configs[0]):
      static       nothing touched since the previous frame (e.g. only the camera moves)
      all_moving   every entity has a velocity in x / y and a tick runs before every frame: all matrices change,
                   no sort key does (+12 B velocity read, +12 B translation write-back per entity)
      keys_change  the frame recomputes everything (key pass, sort, emit: sorted_cache = static_buckets = 0), which is
                   what a frame whose sort keys changed costs; configs[3] also ticks, as its workload says
    frac = entities x algorithmic bytes / frame time / measured HBM peak."""
    stream = torch.cuda.current_stream()
    specs = [
        ("configs[0]", "10k flat sprites, 1 texture, no hierarchy", scenes.config1, 208),
        ("configs[1]", "1M flat sprites, 64 textures x 16 layers x 16 depths", scenes.config2, 208),
        ("configs[2]", "4M-entity scene graph, depth 8: propagation + batching", scenes.config3, 212),
        ("configs[3]", "16M particles: velocity system + transform + batching", scenes.config4, 232),
        ("configs[4]", "64M flat sprites (64 textures x 16 layers x 16 depths) on ONE GPU", scenes.config5, 208),
    ]
    table = []
    for name, what, make, alg in specs:
        if which is not None and name not in which:
            continue
        sc = make()
        native_tick = sc.velocities is not None
        ctx = capi.Context(sc.n, device=device)
        ctx.set_stream(stream.cuda_stream)
        ctx.set_entity_count(sc.n)
        ctx.upload_transforms(0, sc.transforms)
        ctx.upload_sprites(0, sc.sprites)
        if sc.parents is not None:
            ctx.upload_parents(0, sc.parents)
        vel = sc.velocities if native_tick else scenes.config4(n=sc.n_total, first=sc.first, count=sc.n, seed=77).velocities
        dt = 0.01
        if native_tick:
            ctx.upload_velocities(0, vel)

        def run(tick, k):
            for _ in range(k):
                if tick:
                    ctx.system_integrate(dt)
                info = ctx.frame()
            return info

        def timed(tick, bytes_per_entity):
            run(tick, 4)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            info = run(tick, steps)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            assert info.num_instances == sc.n
            return {"ms_per_frame": ms, "entities_per_s": sc.n / (ms * 1e-3), "algorithmic_bytes_per_entity": bytes_per_entity,
                    "frac": sc.n * bytes_per_entity / (ms * 1e-3) / 1e9 / peak, "kernels_per_frame": int(info.kernel_launches),
                    "key_bits": int(info.key_bits), "sort_passes": int(info.sort_passes)}

        base = alg - 24 if native_tick else alg
        modes = {"static": timed(False, base)}
        ctx.set_option("sorted_cache", 0)
        ctx.set_option("static_buckets", 0)
        modes["keys_change"] = timed(native_tick, alg if native_tick else base)
        ctx.set_option("sorted_cache", 1)
        ctx.set_option("static_buckets", 1)
        if not native_tick:
            ctx.upload_velocities(0, vel)
        modes["all_moving"] = timed(True, base + 24)
        info = ctx.frame()
        table.append({"config": name, "workload": what, "entities": sc.n, "batches": int(info.num_batches),
                      "hierarchy_levels": int(info.hierarchy_levels), "modes": modes})
        ctx.close()
        del sc, vel
    return table


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded: ~1.5 M entities/s -> about 0.7 s per step at 2^20
    n_sample = 1 << 20
    cb, step_s = cpu_reference(args.steps, args.warmup, n_sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(N_PER_GPU, args.gpus),
        "note": "reference arm: the Rust reference cannot be built here (no cargo/rustc); this is the CPU oracle port of its "
                "path, single thread like the reference (RefCell storage, sequential systems). Each step times a bounded "
                f"sample — the first {n_sample} entities of the workload scene — and `value` is entities of the sample per "
                "second of CPU time (the path is linear in the entity count)",
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--entities", type=int, default=N_PER_GPU, help="entities per GPU (default: the named workload)")
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--e2e-lanes", type=int, default=3, help="steps in flight in the e2e leg (one context + stream + host thread each)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--option", action="append", default=[], help="tb_ctx_set_option name=value")
    ap.add_argument("--configs", default="all", help="the per-config table at N=1: all | none | comma list, e.g. configs[1],configs[2]")
    ap.add_argument("--strong-entities", type=int, default=1 << 26, help="N > 1: total entities of the strong-scaled configs[4] leg")
    ap.add_argument("--only-configs", action="store_true", help="print just the per-config table (tuning runs)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from tuber_legacy_b200 import capi, scenes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.only_configs:
        which = None if args.configs == "all" else set(args.configs.split(","))
        print(json.dumps({"configs": config_table(capi, scenes, torch, local_rank, measured_peaks()[0], args.steps, which)}))
        return 0
    n = args.entities
    first = rank * n
    sc = scenes.config4(n=world * n, first=first, count=n)
    options = dict(kv.split("=") for kv in args.option)
    ctx = capi.Context(n, device=local_rank, **{k: int(v) for k, v in options.items()})
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.shard_set(rank, world, first)
    ctx.set_entity_count(n)

    # pinned host copies of the component arrays and of the outputs (the e2e leg's buffers)
    def pinned(a):
        buf, ptr = capi.host_alloc(a.nbytes)
        out = buf.view(a.dtype).reshape(a.shape)
        out[...] = a
        return out, ptr

    h_t, p_t = pinned(sc.transforms)
    h_s, p_s = pinned(sc.sprites)
    h_v, p_v = pinned(sc.velocities)
    del sc.transforms, sc.sprites, sc.velocities
    ctx.upload_transforms(0, h_t)
    ctx.upload_sprites(0, h_s)
    ctx.upload_velocities(0, h_v)
    dt = sc.dt

    def step():
        ctx.system_integrate(dt)
        return ctx.frame()

    # ---- device-resident measurement ---------------------------------------------------------
    for _ in range(args.warmup):
        info = step()
    assert info.num_instances == n and info.replanned == 0
    # region A — `value`: exactly K steps between two CUDA events
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        info = step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = ctx.launch_count() - launches0
    assert info.num_instances == n
    # region B — the roofline's kernel durations: the same K steps again with a CUDA-event pair around
    # every kernel launch (tb_profile_*). Kept out of region A because the extra event records cost
    # ~2 % of a 0.68 ms step; region B's own step time is reported beside the kernel times.
    ctx.profile_enable(True)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    k0.record(stream)
    for _ in range(args.steps):
        info = step()
    k1.record(stream)
    barrier()
    ms_total_events = k0.elapsed_time(k1)
    prof = ctx.profile()
    ctx.profile_enable(False)
    assert info.num_instances == n

    # ---- the same steps with nothing reused between frames --------------------------------------
    # (static_buckets=0: key pass + tile offsets + emit every frame; reported beside `value`)
    full = None
    if info.kernel_launches == 1:
        ctx.set_option("static_buckets", 0)
        for _ in range(3):
            step()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        h0.record(stream)
        for _ in range(args.steps):
            finfo = step()
        h1.record(stream)
        barrier()
        full_ms = h0.elapsed_time(h1)
        if world > 1:
            t_ = torch.tensor([full_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            full_ms = float(t_[0])
        full = {"value": n * world * args.steps / (full_ms * 1e-3), "unit": UNIT, "ms_per_step": full_ms / args.steps,
                "kernels_per_frame": int(finfo.kernel_launches),
                "frame_roofline_frac": n * ALG_BYTES_FRAME / (full_ms / args.steps * 1e-3) / 1e9 / measured_peaks()[0],
                "what": "static_buckets=0: nothing cached between frames (key pass, bucket counts, tile offsets and "
                        "emit every frame)"}
        ctx.set_option("static_buckets", 1)
        for _ in range(2):
            step()

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    # Every step: H2D of all three component arrays from pinned memory, tick + frame, D2H of instances,
    # vertices, order and the batch table. PCIe is full duplex, so the steps are double-buffered the way a
    # streaming user would: two contexts (each with its own stream and pinned output buffers) take alternate
    # steps from two host threads, and step k's download overlaps step k+1's upload. `e2e_serial` is the
    # same loop on one context with nothing overlapped.
    class E2E:
        def __init__(self, c):
            self.ctx = c
            hi, self.p_i = capi.host_alloc(n * 80)
            hv, self.p_x = capi.host_alloc(n * 64)
            ho, self.p_o = capi.host_alloc(n * 4)
            self.inst = hi.view(capi.INSTANCE_DTYPE)
            self.vtx = hv.view(capi.VERTEX_DTYPE).reshape(n, 4)
            self.order = ho.view(np.uint32)
            self.out = None

        def step(self):
            c = self.ctx
            c.upload_transforms(0, h_t)
            c.upload_sprites(0, h_s)
            c.upload_velocities(0, h_v)
            c.system_integrate(dt)
            self.out = c.frame()
            c.download_instances(self.out.num_instances, out=self.inst)
            c.download_vertices(self.out.num_instances, out=self.vtx)
            c.download_order(self.out.num_instances, out=self.order)
            self.batches = c.download_batches(self.out.num_batches)

        def free(self):
            for p_ in (self.p_i, self.p_x, self.p_o):
                capi.host_free(p_)

    ctx.set_stream(0)  # e2e contexts run on their own streams
    nlanes = max(2, args.e2e_lanes)
    extra = []
    for _ in range(nlanes - 1):
        cx = capi.Context(n, device=local_rank, **{k: int(v) for k, v in options.items()})
        cx.shard_set(rank, world, first)
        cx.set_entity_count(n)
        extra.append(cx)
    lanes = [E2E(ctx)] + [E2E(cx) for cx in extra]
    for ln in lanes:
        ln.step()
    # serial: one context, nothing overlapped
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(2, args.e2e_steps // 2)):
        lanes[0].step()
    torch.cuda.synchronize()
    serial_ms = (time.perf_counter() - t0) * 1e3 / max(2, args.e2e_steps // 2)
    # pipelined: one step in flight per lane
    per_lane = max(1, -(-args.e2e_steps // nlanes))
    e2e_total = nlanes * per_lane

    def lane_loop(ln, k, delay):
        torch.cuda.set_device(local_rank)
        time.sleep(delay)  # out of phase: the second lane uploads while the first one downloads
        for _ in range(k):
            ln.step()

    barrier()
    t0 = time.perf_counter()
    threads = [threading.Thread(target=lane_loop, args=(ln, per_lane, i * serial_ms * 1e-3 / nlanes))
               for i, ln in enumerate(lanes)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    out = lanes[0].out
    order = lanes[0].order
    assert lanes[1].out.num_instances == out.num_instances and np.array_equal(lanes[1].order[:4096], order[:4096])
    h2d = h_t.nbytes + h_s.nbytes + h_v.nbytes
    d2h = int(out.num_instances) * (80 + 64 + 4) + int(out.num_batches) * 16
    checksum = int(order[:: max(1, n // 4096)].astype(np.uint64).sum())
    for cx in extra:
        cx.close()
    for ln in lanes:
        ln.free()
    ctx.set_stream(stream.cuda_stream)

    # ---- N > 1: the same scene as ONE draw list on the presenting rank -------------------------------
    # Exact global draw order by key-histogram placement (two small NCCL collectives per frame) with
    # every rank's emit kernel storing its instances straight into rank 0's buffers over NVLink
    # (CUDA IPC peer mapping): the payload never goes through a collective. Reported next to `value`.
    gather = None
    strong = None
    if world > 1:
        def dev_tensor(ptr, nbytes):
            class _Dev:
                pass
            d = _Dev()
            d.__cuda_array_interface__ = {"shape": (int(nbytes) // 8,), "typestr": "<i8", "data": (int(ptr), False), "version": 3}
            return torch.as_tensor(d, device=torch.device("cuda", local_rank))

        def share_id():
            box = [capi.shard_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            return box[0]

        def timed_sharded(c, tick, steps_):
            for _ in range(2):
                if tick:
                    c.system_integrate(dt)
                o = c.frame_sharded()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t_ = time.perf_counter()
            a0.record(stream)
            for _ in range(steps_):
                if tick:
                    c.system_integrate(dt)
                o = c.frame_sharded()
            a1.record(stream)
            barrier()
            ms_ = max(a0.elapsed_time(a1), (time.perf_counter() - t_) * 1e3) / steps_
            tt = torch.tensor([ms_], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt[0]), o

        method = ("tb_shard_init + tb_frame_sharded: the exchanges run inside the library (NCCL opened by the library). The first "
                  "frame places the list (3-word and short-key-histogram all-gathers, one barrier); the timed frames are steady "
                  "ones: one integer all-reduce of agreement, then every rank's emit kernels (velocity system fused, keys "
                  "re-checked) store a 64-byte world row and the entity id per element (68 B instead of the 84 B of an instance) at "
                  "their kept positions of the exact global draw order straight into rank 0's memory over NVLink (CUDA IPC "
                  "mapping) in 16 launches, each followed by completion flags; rank 0 expands the rows into instances and quad "
                  "vertices behind those flags (same arithmetic: bit-identical); a closing all-reduce carries every rank's verdict")
        # (1) weak: the headline workload, 2^24 entities per rank, as ONE draw list on rank 0
        n_total = n * world
        ctx.shard_init(rank, world, share_id(), first, 0)
        gsteps = max(5, args.steps // 2)
        gather_ms, gout = timed_sharded(ctx, True, gsteps)
        assert gout.num_instances == n_total
        remote_bytes = (world - 1) * n * (64 + 4)  # steady frames ship a 64-byte world row + the entity id per element
        gcheck = None
        if rank == 0:
            # integrity of the list the timed frames left on rank 0 (a steady frame: kept placement, vertices expanded behind
            # the other ranks' completion flags): every entity appears exactly once, and the list's vertices are exactly
            # what a fresh expansion of the list's instances gives
            od = dev_tensor(gout.d_order, n_total * 4).view(torch.int32)
            perm = bool(torch.equal(torch.sort(od).values, torch.arange(n_total, dtype=torch.int32, device=od.device)))
            del od
            scratch = capi.device_alloc(n_total * 64, local_rank)
            ctx.vertices_from_instances(gout.d_instances, 0, n_total, scratch)
            vsame = bool(torch.equal(dev_tensor(scratch, n_total * 64), dev_tensor(gout.d_vertices, n_total * 64)))
            capi.device_free(scratch)
            gcheck = {"order_is_a_permutation_of_all_entities": perm, "vertices_equal_a_fresh_expansion_of_the_instances": vsame,
                      "steady_frame": int(gout.sort_passes) == 0}
        gather = {"value": n_total / (gather_ms * 1e-3), "unit": UNIT, "ms_per_step": gather_ms, "steps": gsteps,
                  "global_instances": n_total, "bytes_into_rank0_per_step": remote_bytes,
                  "rank0_ingress_GBps": remote_bytes / (gather_ms * 1e-3) / 1e9, "scaling": "weak", "check": gcheck, "method": method}
        ctx.shard_finalize()

        # (2) strong: configs[4] as BASELINE names it — 2^26 entities in total, id-range-partitioned over the ranks
        n5 = args.strong_entities
        f5, c5 = scenes.shard_range(n5, rank, world)
        s5 = scenes.config5(n=n5, first=f5, count=c5)
        c5x = capi.Context(c5, device=local_rank)
        c5x.set_stream(stream.cuda_stream)
        c5x.shard_init(rank, world, share_id(), f5, 0)
        c5x.set_entity_count(c5)
        c5x.upload_transforms(0, s5.transforms)
        c5x.upload_sprites(0, s5.sprites)
        del s5

        def timed_local(k):
            for _ in range(3):
                o = c5x.frame()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a0.record(stream)
            for _ in range(k):
                o = c5x.frame()
            a1.record(stream)
            barrier()
            tt = torch.tensor([a0.elapsed_time(a1) / k], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt[0]), o

        c5x.set_output(0, 0, 0, 0)
        static_ms, _ = timed_local(gsteps)                      # every rank frames its own shard, nothing exchanged
        c5x.set_option("sorted_cache", 0)
        c5x.set_option("static_buckets", 0)
        full_ms, _ = timed_local(gsteps)
        c5x.set_option("sorted_cache", 1)
        c5x.set_option("static_buckets", 1)
        list_ms, lout = timed_sharded(c5x, False, gsteps)       # one draw list on rank 0
        assert lout.num_instances == n5
        # where a list frame's time goes on every rank: per-kernel CUDA-event times (a second pass, not the timed one)
        c5x.profile_enable(True)
        t_wall = time.perf_counter()
        for _ in range(3):
            c5x.frame_sharded()
        wall_ms = (time.perf_counter() - t_wall) * 1e3 / 3
        mine = {"rank": rank, "wall_ms_per_frame": round(wall_ms, 3),
                "kernels_ms_per_frame": {k_: round(v_[0] / 3, 4) for k_, v_ in c5x.profile().items()}}
        c5x.profile_enable(False)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
        check = None
        if rank == 0:
            # the presenting rank's global list against a single-context frame of the whole scene, byte for byte
            whole = scenes.config5(n=n5)
            one = capi.Context(n5, device=local_rank)
            one.set_stream(stream.cuda_stream)
            one.set_entity_count(n5)
            one.upload_transforms(0, whole.transforms)
            one.upload_sprites(0, whole.sprites)
            del whole
            oo = one.frame()
            m5 = int(lout.num_instances)
            eq = {"order": bool(torch.equal(dev_tensor(lout.d_order, m5 * 4), dev_tensor(oo.d_order, m5 * 4))),
                  "instances": bool(torch.equal(dev_tensor(lout.d_instances, m5 * 80), dev_tensor(oo.d_instances, m5 * 80))),
                  "vertices": bool(torch.equal(dev_tensor(lout.d_vertices, m5 * 64), dev_tensor(oo.d_vertices, m5 * 64))),
                  "batches": bool(int(oo.num_batches) == int(lout.num_batches) and
                                  c5x.download_batches(lout.num_batches).tobytes() == one.download_batches(oo.num_batches).tobytes())}
            check = {"equal_to_single_gpu_frame": all(eq.values()), **eq, "compared_on": "device (torch.equal over the raw buffers)"}
            one.close()
        dist.barrier()
        rb5 = (n5 - c5 if rank == 0 else 0)
        rb5t = torch.tensor([float(rb5)], dtype=torch.float64, device="cuda")
        dist.all_reduce(rb5t, op=dist.ReduceOp.MAX)
        remote5 = float(rb5t[0]) * (64 + 4)
        strong = {"workload": f"configs[4]: {n5} flat sprites in total (64 textures x 16 layers x 16 depths), contiguous id ranges over "
                              f"{world} GPUs, instance list gathered on rank 0", "scaling": "strong", "entities_total": n5,
                  "compute_only": {"static": {"ms_per_frame": static_ms, "entities_per_s": n5 / (static_ms * 1e-3)},
                                   "keys_change": {"ms_per_frame": full_ms, "entities_per_s": n5 / (full_ms * 1e-3)},
                                   "what": "every rank frames its own shard into its own buffers; max over ranks"},
                  "list_on_rank0": {"ms_per_frame": list_ms, "entities_per_s": n5 / (list_ms * 1e-3),
                                    "bytes_into_rank0_per_frame": remote5, "rank0_ingress_GBps": remote5 / (list_ms * 1e-3) / 1e9,
                                    "what": "end to end, transfer overlapped with the emit: the remote stores ARE the emit kernels' stores; "
                                            "includes the agreement, the closing all-reduce and the vertex regeneration on rank 0",
                                    "per_rank": per_rank},
                  "check": check, "method": method}
        c5x.shard_finalize()
        c5x.close()

    # ---- reduce over ranks: max time, summed work -------------------------------------------------
    vals = torch.tensor([ms_total, e2e_ms, float(gpu_launches), serial_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_ms, gpu_launches, serial_ms = float(mx[0]), float(mx[1]), float(sm[2]), float(mx[3])
    total_entities = n * world

    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_per_step = ms_total / args.steps
        value = total_entities * args.steps / (ms_total * 1e-3)
        emit_ms, _, emit_launches = prof.get("emit", (0.0, 0, 0))
        emit_ms_avg = emit_ms / max(emit_launches, 1)
        # steady-state frames of this workload are ONE kernel (velocity system + rank + compose + emit):
        # then the dominant kernel's algorithmic bytes are the whole frame's
        fused = info.kernel_launches == 1
        alg_emit = ALG_BYTES_FRAME if fused else ALG_BYTES_EMIT
        achieved = n * alg_emit / (emit_ms_avg * 1e-3) / 1e9 if emit_ms_avg > 0 else None
        kernels = {k: {"ms_per_launch": v[0] / max(v[2], 1), "launches_per_step": v[2] / args.steps,
                       "share_of_step": v[0] / ms_total_events} for k, v in prof.items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(n, world),
            "frame": {"path": {1: "bucket (one fused rank+compose+emit pass)", 2: "sort"}.get(info.path),
                      "key_bits": info.key_bits, "sort_passes": info.sort_passes, "kernels_per_frame": info.kernel_launches,
                      "placement": "per-tile bucket offsets cached from the last full frame, validated every frame "
                                   "(buckets depend on the Sprite only)" if info.kernel_launches == 1 else "recomputed every frame",
                      "options": options},
            "clocks": clocks,
            "e2e": {"value": total_entities * e2e_total / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / e2e_total,
                    "steps": e2e_total, "steps_in_flight": nlanes, "host_buffers": "pinned (tb_host_alloc)", "rank0_numa": numa,
                    "h2d_GBps_per_rank": h2d / (e2e_ms / e2e_total * 1e-3) / 1e9, "d2h_GBps_per_rank": d2h / (e2e_ms / e2e_total * 1e-3) / 1e9,
                    "how": "contexts / streams / host threads take alternate steps so one step's D2H overlaps the next steps' "
                           "H2D (PCIe full duplex); every step still copies all inputs in and all outputs out. The library lets "
                           "one bulk copy per direction onto the link at a time and reads its small results back without a copy "
                           "engine; the timed region includes the first step's upload and the last step's download, which "
                           "overlap nothing",
                    "link": "this box, same byte counts, plain pinned copies on two streams: H2D alone 55.5 GB/s, D2H alone "
                            "57.2 GB/s, both directions together 48.9 ms = 76.9 GB/s (tools/microbench_pcie.py)",
                    "order_checksum": checksum},
            "e2e_serial": {"value": total_entities / (serial_ms * 1e-3), "unit": UNIT, "ms_per_step": serial_ms,
                           "steps_in_flight": 1},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "kernel": "emit_small_kernel<uint32_t, %s>" % ("true" if fused else "false"),
                         "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_entity": alg_emit,
                         "ms_per_launch": emit_ms_avg,
                         "timing": "CUDA-event pair around every launch over a second pass of the same K steps "
                                   "(%.4f ms per step with those events; `value` is timed without them)" % (ms_total_events / args.steps),
                         "kernel_does": "velocity system + key statistics + rank + compose + instance/vertex emission (single-kernel frame)"
                         if fused else "rank + compose + instance/vertex emission"},
            "frame_roofline": {"algorithmic_bytes_per_entity": ALG_BYTES_FRAME,
                               "achieved": n * ALG_BYTES_FRAME / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": n * ALG_BYTES_FRAME / (ms_per_step * 1e-3) / 1e9 / peak,
                               "note": "whole step (all kernels + the per-frame host sync) against the same peak"},
            "kernels": kernels,
        }
        # DRAM bytes per launch of the dominant kernel from an `ncu --set full` capture — reported only when that capture
        # was taken from THIS build (same source hash) or from a build with the very same dominant kernel (same SASS hash),
        # otherwise null
        traffic_file = os.path.join(ROOT, "profiles", "emit_traffic.json")
        if os.path.exists(traffic_file):
            try:
                from tuber_legacy_b200 import build as tb_build
                tr = json.load(open(traffic_file))
                same_sources = tr.get("source_sha256") == tb_build.source_hash()
                same_kernel = False
                if not same_sources and tr.get("kernel_sass_sha256"):
                    try:  # another build, but is the dominant kernel instruction for instruction the profiled one?
                        same_kernel = tr["kernel_sass_sha256"] == tb_build.kernel_sass_hash()
                    except Exception:
                        same_kernel = False
                if (same_sources or same_kernel) and tr.get("entities") == n:
                    line["roofline"]["traffic"] = tr.get("dram_bytes_per_launch")
                    line["roofline"]["traffic_source"] = tr.get("source") + ("; same sources as this build" if same_sources else
                                                                              "; taken from another build whose dominant kernel has the same SASS as this "
                                                                              "build's (sha256 over the instruction text, constant-bank offsets normalised)")
                else:
                    line["roofline"]["traffic_source"] = "profiles/emit_traffic.json is from another build: not reported"
            except Exception:
                pass
        if full is not None:
            line["full_recompute_every_frame"] = full
        if gather is not None:
            line["gather_to_rank0"] = gather
        if strong is not None:
            line["strong_scaling_configs4"] = strong
        if world == 1 and not args.no_cpu:
            cb, _ = cpu_reference(3, 1, CPU_SAMPLE)
            line["cpu_baseline"] = cb
    ctx.close()
    for p in (p_t, p_s, p_v):
        capi.host_free(p)
    if rank == 0:
        if world == 1 and args.configs != "none":
            which = None if args.configs == "all" else set(args.configs.split(","))
            line["configs"] = config_table(capi, scenes, torch, local_rank, peak, min(args.steps, 20), which)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
