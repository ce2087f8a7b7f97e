#!/usr/bin/env python
"""bench.py — Msamples/s of the wavefront path tracer on BASELINE.json's configuration.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                      (the reference's own CPU path on the host cores)

A "step" is one full render of the workload (cbox 1024x1024, 256 spp, max_depth 6 — BASELINE
configs[1]) through the hot path. `value` times the render with scene + queues resident in HBM
and the film left on the device (CUDA events on the launching stream, max over ranks, incl. the
final NCCL reduce of the film for N > 1); `e2e` times the public call PathIntegrator.render()
with a HOST film buffer (per-step host->device work-item tables and device->host film copy
inside the timed region). Pixel blocks are sharded over ranks (block b -> rank b % N); seeds
are a pure function of (block, pixel, pass) so the image does not depend on N.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# SURVEY.md §8(d) algorithmic bytes (per unit, read + write, excluding BVH / geometry)
B_RAYGEN, B_HIT, B_SHADE, B_SHADOW, B_COMPACT, B_FILM = 80, 52, 224, 68, 8, 40


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2pt", choices=["b2pt", "reference"])
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--spp", type=int, default=256)
    ap.add_argument("--spp-per-pass", type=int, default=16,
                    help="integrator 'samples_per_pass' (integrator.cpp:34); 0 = spp")
    ap.add_argument("--max-depth", type=int, default=6)
    ap.add_argument("--lane-slots", type=int, default=0)
    ap.add_argument("--rfilter", default=None, choices=["box", "gaussian"], help="matball: override the reconstruction filter")
    ap.add_argument("--cpu-baseline", type=int, default=1)
    ap.add_argument("--no-stage-timing", action="store_true")
    ap.add_argument("--workload", default="cbox", choices=["cbox", "veach", "grid", "matball"],
                    help="cbox = BASELINE configs[1] (the bench line); veach / grid = configs[3] / configs[4] (extra measurements)")
    ap.add_argument("--single-set", action="store_true", help="one lane set on one stream (A/B of the two-stream overlap)")
    ap.add_argument("--persistent-trace", action="store_true", help="B2_FLAG_PERSISTENT_TRACE: persistent-warp traversal kernel also without instances (A/B)")
    ap.add_argument("--resume-trace", action="store_true", help="B2_FLAG_RESUME_TRACE: yield / resume traversal kernel instead of the static ray -> lane kernel (A/B)")
    ap.add_argument("--trace-block", type=int, default=0, choices=[0, 256, 128, 64], help="threads per CTA of the traversal kernel (0 = the library's choice)")
    ap.add_argument("--shade-block", type=int, default=0, choices=[0, 256, 128, 64], help="threads per CTA of the shading kernel (0 = the library's choice)")
    ap.add_argument("--yield-threshold", type=int, default=0, help="k_trace_resume: suspend below this many traversing lanes (0 = default)")
    ap.add_argument("--bvh-counters", action="store_true", help="cbox: also run the BVH-counting step (always on for the other workloads)")
    ap.add_argument("--parity", action="store_true", help="cbox: also render the configuration with the reference CLI and report rel-L2")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# reference CPU arm (oracle/_ref = the unmodified reference CLI; falls back to the oracle port)
# ------------------------------------------------------------------------------------------------
def _ref_dist():
    d = os.path.join(ROOT, "oracle", "_ref", "dist")
    return d if os.path.exists(os.path.join(d, "mitsuba")) else None


def _ref_cbox_xml(res, spp, spp_pass, max_depth):
    mesh_dir = os.path.join(ROOT, "oracle", "_ref", "scenes", "cbox", "meshes")
    shapes = [("cbox_luminaire.obj", "light", True), ("cbox_floor.obj", "white", False), ("cbox_ceiling.obj", "white", False),
              ("cbox_back.obj", "white", False), ("cbox_greenwall.obj", "green", False), ("cbox_redwall.obj", "red", False),
              ("cbox_smallbox.obj", "box", False), ("cbox_largebox.obj", "box", False)]
    from mitsuba2_b200.scenes import CBOX_BSDFS
    bs = "".join(f'<bsdf type="diffuse" id="{k}"><rgb name="reflectance" value="{v[0]}, {v[1]}, {v[2]}"/></bsdf>'
                 for k, v in CBOX_BSDFS.items())
    sh = ""
    for fn, b, lum in shapes:
        t = '<transform name="to_world"><translate x="0" y="-0.5" z="0"/></transform>' if lum else ""
        em = '<emitter type="area"><rgb name="radiance" value="18.387, 10.9873, 2.75357"/></emitter>' if lum else ""
        sh += f'<shape type="obj"><string name="filename" value="{mesh_dir}/{fn}"/>{t}<ref id="{b}"/>{em}</shape>'
    sp = f'<integer name="samples_per_pass" value="{spp_pass}"/>' if spp_pass and spp_pass < spp else ""
    return f"""<scene version="2.0.0"><integrator type="path"><integer name="max_depth" value="{max_depth}"/>
<integer name="block_size" value="32"/>{sp}</integrator>
<sensor type="perspective"><string name="fov_axis" value="smaller"/><float name="near_clip" value="10"/>
<float name="far_clip" value="2800"/><float name="fov" value="39.3077"/>
<transform name="to_world"><lookat origin="278, 273, -800" target="278, 273, -799" up="0, 1, 0"/></transform>
<sampler type="independent"><integer name="sample_count" value="{spp}"/></sampler>
<film type="hdrfilm"><integer name="width" value="{res}"/><integer name="height" value="{res}"/><rfilter type="box"/>
<string name="pixel_format" value="rgb"/><string name="component_format" value="float32"/><string name="file_format" value="pfm"/></film>
</sensor>{bs}{sh}</scene>"""


def run_reference_cli(res, spp, spp_pass, max_depth, threads):
    """-> (seconds of the render as logged by the reference, samples) or None if it cannot run here."""
    dist = _ref_dist()
    if not dist:
        return None
    env = dict(os.environ); env["LD_LIBRARY_PATH"] = dist + ":" + env.get("LD_LIBRARY_PATH", "")
    with tempfile.TemporaryDirectory() as td:
        sp = os.path.join(td, "s.xml")
        open(sp, "w").write(_ref_cbox_xml(res, spp, spp_pass, max_depth))
        try:
            r = subprocess.run([os.path.join(dist, "mitsuba"), "-m", "scalar_rgb", "-t", str(threads), "-o",
                                os.path.join(td, "o.pfm"), sp], env=env, capture_output=True, text=True, timeout=1200)
        except Exception:
            return None
        if r.returncode != 0:
            return None
        # "Rendering finished. (took 1.23s)"  integrator.cpp:174-176
        import re
        m = re.search(r"Rendering finished\. \(took ([0-9.]+)(ms|s|m)", r.stdout + r.stderr)
        if not m:
            return None
        t = float(m.group(1)) * {"ms": 1e-3, "s": 1.0, "m": 60.0}[m.group(2)]
        return t, res * res * spp


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of `kernel`, from the committed
    `ncu --set full` capture of the final build: profiles/r02_ncu_<workload>.csv with rows
    kernel,duration_us,dram_read_bytes,dram_write_bytes,threads_per_warp_instr,issue_active_pct (written by
    scripts/ncu_summary.py from the .ncu-rep). None when no capture is committed for this workload."""
    path = os.path.join(ROOT, "profiles", f"r02_ncu_{workload}.csv")
    try:
        rows = [ln.strip().split(",") for ln in open(path) if ln.strip() and not ln.startswith("#")]
        hdr = rows[0]
        for r in rows[1:]:
            d = dict(zip(hdr, r))
            if d["kernel"].startswith(kernel):
                def f(k):
                    try:
                        return float(d[k])
                    except Exception:
                        return None
                return {"bytes": float(d["dram_read_bytes"]) + float(d["dram_write_bytes"]), "duration_us": float(d["duration_us"]),
                        "issue_active_pct": f("issue_active_pct"), "lanes_per_warp_instr": f("threads_per_warp_instr"),
                        "source": os.path.relpath(path, ROOT)}
    except Exception:
        pass
    return None


def run_oracle_port(res, spp, spp_pass, max_depth, threads):
    from mitsuba2_b200 import abi, scenes
    from tests import oracle_lib as O
    sb = scenes.cbox(O.OracleHost(), res, res)
    rp = abi.render_params(max_depth=max_depth, rr_depth=5, spp=spp, samples_per_pass=spp_pass, block_size=32)
    t0 = time.perf_counter()
    O.render(sb, rp, threads=threads)
    return time.perf_counter() - t0, res * res * spp


def usable_cores():
    """Host threads this process may actually use: the affinity mask, capped by the cgroup CPU quota (a leased box
    reports every core of the node in os.cpu_count() even when the container may only burn a few of them)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0]); per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return max(1, n)


def cpu_sample(a, spp):
    """One bounded CPU sample of the workload: same scene / resolution / depth, fewer spp."""
    cores = usable_cores()
    spp_pass = min(a.spp_per_pass or spp, spp)
    r = run_reference_cli(a.res, spp, spp_pass, a.max_depth, cores)
    kind = "reference"
    if r is None:
        r = run_oracle_port(a.res, spp, spp_pass, a.max_depth, cores)
        kind = "port"
    t, n = r
    return dict(value=n / t / 1e6, unit="Msamples/s", cores=cores, kind=kind,
                sample=f"cbox {a.res}x{a.res}, {spp} spp (samples_per_pass {spp_pass}), max_depth {a.max_depth}, "
                       f"{'mitsuba -m scalar_rgb -t %d' % cores if kind == 'reference' else 'oracle/b2o.c, %d threads' % cores}, {t:.2f} s")


def main_reference(a, rank):
    if rank != 0:
        return
    spp = 16                      # bounded sample per step (~4 s of host time on 128 threads)
    warm = min(a.warmup, 1)       # one warm-up sample pages the binary and the scene in; more would only burn minutes
    for _ in range(warm):
        cpu_sample(a, spp)
    vals = []
    t0 = time.perf_counter()
    for _ in range(a.steps):
        vals.append(cpu_sample(a, spp))
    dt = time.perf_counter() - t0
    v = float(np.mean([x["value"] for x in vals]))
    cb = dict(vals[-1]); cb["value"] = v
    emit(json.dumps({
        "impl": "reference", "metric": "Msamples/s", "value": v, "unit": "Msamples/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": warm, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cbox {a.res}x{a.res}, {a.spp} spp (timed on a {spp}-spp sample per step), max_depth {a.max_depth}, "
                               "diffuse + area emitter, box filter, independent sampler seed 0, block_size 32"},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index; self.rows = []; self.p = None; self.t0 = None; self.t1 = None

    def mark(self, begin):
        """bracket the timed region: samples inside it are the ones reported (all samples under load if it was too short)"""
        if begin:
            self.t0 = time.time()
        else:
            self.t1 = time.time()

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")] + [time.time()])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        rows = [r for r in self.rows if len(r) >= 10 and r[1].replace(".", "").isdigit()]
        timed = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[-1] <= self.t1 + 0.05]
        window = "timed region"
        if len(timed) < 3:     # a very short timed region: every sample since the warm-up started ran under the same load
            timed = rows; window = "warm-up + timed region"
        sm = [float(r[1]) for r in timed]
        mx = [float(r[2]) for r in timed if r[2].replace(".", "").isdigit()]
        reasons = set()
        self.window = window
        for r in timed:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower() == "active":
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


from mitsuba2_b200.dist import CudaArray  # noqa: E402,F401


def main_b2pt(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from mitsuba2_b200 import abi, api, scenes

    if not torch.cuda.is_available() or api.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (libb2pt has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if a.workload == "veach":
        sb = scenes.veach_mis(api.Host(), a.res, a.res * 9 // 16)
    elif a.workload == "grid":
        sb = scenes.instanced_grid(api.Host(), a.res, a.res)
    elif a.workload == "matball":
        sb = scenes.matball(api.Host(), a.res, a.res * 9 // 16, rfilter=a.rfilter or "gaussian")
    else:
        sb = scenes.cbox(api.Host(), a.res, a.res)
    t0 = time.perf_counter()
    scene = api.Scene(sb)
    t_scene = time.perf_counter() - t0
    integ = api.PathIntegrator(max_depth=a.max_depth, rr_depth=5, samples_per_pass=a.spp_per_pass, block_size=32)
    H, W = sb.film.crop_size[1], sb.film.crop_size[0]
    n_samples = H * W * a.spp
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")     # > 126 MB L2
    host_film = torch.empty((H, W, 5), dtype=torch.float32).pin_memory().numpy()
    base_flags = ((abi.FLAG_SINGLE_SET if a.single_set else 0) | (abi.FLAG_PERSISTENT_TRACE if a.persistent_trace else 0) |
                  (abi.FLAG_RESUME_TRACE if a.resume_trace else 0) | abi.flag_yield_threshold(a.yield_threshold) |
                  abi.flag_blocks({0: 0, 256: 1, 128: 2, 64: 3}[a.trace_block], {0: 0, 256: 1, 128: 2, 64: 3}[a.shade_block]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_film(ptr, n, stream):
        """the single collective of the job: XYZAW film of every rank -> rank 0 (NCCL, on the device)"""
        film = torch.as_tensor(CudaArray(ptr, (n,)), device="cuda")
        dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)

    def step_device(flags=0):
        """-> device ms of this rank's render (+ film reduce), stats"""
        flush.zero_()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ptr = integ.render_device(scene, a.spp, rank=rank, world_size=world, lane_slots=a.lane_slots, flags=flags | base_flags)
        if world > 1:
            reduce_film(ptr, H * W * 5, None)
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1), dict(integ.stats)

    def step_e2e():
        """the public call with a HOST film: b2_render (1 GPU) / b2_render_reduce (N GPUs: device-side NCCL reduce through
        the callback, one device->host copy on rank 0)"""
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world > 1:
            integ.render_reduce(scene, a.spp, reduce_film, rank=rank, world_size=world, lane_slots=a.lane_slots, out=host_film, flags=base_flags)
            torch.cuda.synchronize()
        else:
            integ.render(scene, a.spp, lane_slots=a.lane_slots, out=host_film, flags=base_flags)
        return time.perf_counter() - t0

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()          # (nvidia-smi needs ~1 s to deliver its first sample: started before the warm-up)
    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()
    clocks.mark(True)
    ms_steps, stats = [], []
    for _ in range(a.steps):
        ms, st = step_device()
        ms_steps.append(ms); stats.append(st)
    barrier()
    clocks.mark(False)
    total_ms = torch.tensor([sum(ms_steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    clk = clocks.stop() if rank == 0 else None

    # e2e: same metric through the public host-buffer API
    step_e2e()
    barrier()
    e2e_t = []
    for _ in range(a.steps):
        e2e_t.append(step_e2e())
    barrier()
    e2e_total = torch.tensor([sum(e2e_t)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    e2e_total = float(e2e_total.item())

    # per-stage device time: extra steps with every launch bracketed by CUDA events (one lane set, so that a stage's
    # time is its own), averaged over >= 3 steps so that `kernel` (the stage with the larger time) is a stable choice
    roof = None; stages = None; bvh = None
    if not a.no_stage_timing:
        n_t = max(3, min(a.steps, 5))
        acc = {"trace": 0.0, "shade": 0.0}; ms_t = 0.0; st = None
        for _ in range(n_t):
            m, st = step_device(flags=abi.FLAG_STAGE_TIMING)
            ms_t += m / n_t; acc["trace"] += st["ms_intersect"] / n_t; acc["shade"] += st["ms_shade"] / n_t
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        n_it = st["n_iterations"]
        # algorithmic bytes per stage. k_trace = K2 52 B/ray + K5 8 B/lane (the ballot compaction happens at the
        # end of the traversal) + K3 68 B/shadow ray; k_shade = K4 224 B/segment + K6 40 B/sample put + K1 80 B/regenerated sample
        stage_bytes = {
            "trace": st["n_segments"] * (B_HIT + B_COMPACT) + st["n_shadow_rays"] * B_SHADOW,
            "shade": st["n_segments"] * B_SHADE + st["n_samples"] * (B_RAYGEN + B_FILM),
        }
        stage_units = {"trace": st["n_segments"] + st["n_shadow_rays"], "shade": st["n_segments"]}
        stages = {}
        for k, nbytes in stage_bytes.items():
            t = acc[k]
            gbs = (nbytes / (t * 1e-3) / 1e9) if t > 0 else None
            tr = ncu_traffic(a.workload, "k_" + k)
            stages[k] = {"kernel": "k_" + k, "ms": t, "units": stage_units[k], "bytes": nbytes, "gbs": gbs, "frac": gbs / peak if gbs else None,
                         "launches": n_it, "avg_launch_us": t / max(n_it, 1) * 1e3, "bytes_per_launch": nbytes / max(n_it, 1),
                         "traffic": tr["bytes"] if tr else None, "traffic_source": tr["source"] if tr else None,
                         # from the same committed ncu capture: what bounds a kernel whose HBM fraction is low — the share of issue
                         # slots it keeps busy and the lanes active per warp instruction (not measured in this run)
                         "ncu_issue_active_pct": tr["issue_active_pct"] if tr else None,
                         "ncu_lanes_per_warp_instr": tr["lanes_per_warp_instr"] if tr else None}
        dom = max(stages, key=lambda k: stages[k]["ms"])
        tot_bytes = sum(stage_bytes.values()); tot_t = sum(acc.values())
        roof = {"bound": "hbm", "kernel": "k_" + dom, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["frac"], "traffic": stages[dom]["traffic"], "traffic_source": stages[dom]["traffic_source"],
                "kernel_rule": f"stage with the larger device time, averaged over {n_t} stage-timed steps",
                "achieved_bytes_per_launch": stages[dom]["bytes_per_launch"],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "stages": stages,
                "all_stages_gbs": tot_bytes / (tot_t * 1e-3) / 1e9 if tot_t > 0 else None,
                "all_stages_frac": tot_bytes / (tot_t * 1e-3) / 1e9 / peak if tot_t > 0 else None,
                "step_ms_with_timing": ms_t,
                "note": ("the wavefront is sized to stay in L2 and the stage kernels are instruction- / latency-bound, not HBM-bound: the "
                         "HBM fraction is what the contract asks for; stages[*].ncu_issue_active_pct / ncu_lanes_per_warp_instr (committed "
                         "ncu capture) say how busy the issue slots are")}
        # BVH term: nodes fetched and primitives tested by all rays of one step, from the counting
        # instantiation of the traversal kernel (slower: its time is not used)
        if a.workload != "cbox" or a.bvh_counters:
            _, sc = step_device(flags=abi.FLAG_COUNT_BVH)
            nb = sc["n_nodes_visited"] * 128 + sc["n_tris_tested"] * 48
            rays = sc["n_segments"] + sc["n_shadow_rays"]
            bvh = {"nodes_visited": sc["n_nodes_visited"], "tris_tested": sc["n_tris_tested"], "rays": rays,
                   "nodes_per_ray": sc["n_nodes_visited"] / max(rays, 1), "tris_per_ray": sc["n_tris_tested"] / max(rays, 1),
                   "node_bytes": 128, "tri_bytes": 48, "bytes": nb,
                   "bytes_at_survey_node_size": sc["n_nodes_visited"] * 80 + sc["n_tris_tested"] * 48,
                   "gbs_over_trace_time": nb / (acc["trace"] * 1e-3) / 1e9 if acc["trace"] > 0 else None,
                   "frac_over_trace_time": nb / (acc["trace"] * 1e-3) / 1e9 / peak if acc["trace"] > 0 else None,
                   "bvh_nodes": sc["bvh_nodes"], "bvh_prims": sc["bvh_tris"], "bvh_depth": sc["bvh_depth"],
                   "note": "requested bytes (mostly served by L1/L2: the tree is smaller than the 126 MB L2)"}

    if rank == 0:
        st = stats[-1]
        value = n_samples * a.steps / (total_ms * 1e-3) / 1e6
        e2e_v = n_samples * a.steps / e2e_total / 1e6
        q_bytes = (scene_lane_slots(a, st) * 172) / 1e6
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": {"cbox": f"cbox {W}x{H}, {a.spp} spp, max_depth {a.max_depth}, diffuse + area emitter, box filter (BASELINE configs[1])",
                                    "veach": f"veach-mis class {W}x{H}, {a.spp} spp, 4 emitters, dielectric + rough conductors (BASELINE configs[3])",
                                    "matball": f"rough-conductor (GGX) ball on a checkerboard plane under an environment map, {sb.n_triangles} triangles, {W}x{H}, {a.spp} spp, gaussian filter (BASELINE configs[2] class)",
                                    "grid": f"instanced displaced grid, {sb.n_triangles} triangles x {len(sb.instances)} instances, {W}x{H}, {a.spp} spp (BASELINE configs[4])"}[a.workload], "samples_per_pass": a.spp_per_pass or a.spp, "block_size": 32, "seed": 0,
                       "parallelism": f"pixel blocks b % {world} == rank, one NCCL reduce of the film" if world > 1 else "1 GPU",
                       "l2": f"256 MB flush (> the 126 MB L2) between steps; within a step the wavefront queues (~{q_bytes:.0f} MB) are "
                             "sized to stay L2-resident between the stages, so per-step DRAM traffic is below the algorithmic bytes",
                       "lane_slots": a.lane_slots or "auto"},
            "mean_segments_per_sample": st["n_segments"] * world / n_samples if world == 1 else None,
            "scene_create_s": t_scene,
            "e2e": {"value": e2e_v, "unit": "Msamples/s",
                    # (the work-item tables are uploaded once per (film, block size, rank, world) and cached in the scene)
                    "h2d_bytes_per_step": int(ctypes_sizeof_params()),
                    "d2h_bytes_per_step": int(H * W * 5 * 4 + 128 * (st["n_iterations"] // 8) + 64),
                    "api": "b2_render" if world == 1 else "b2_render_reduce (device-side NCCL reduce in the callback; rank 0 copies the film)"},
            "gpu_launches": int(sum(s["n_launches"] for s in stats)),
            "clocks": clk,
        }
        if roof:
            # whole-step view: algorithmic bytes of all stages / the timed step (two lane sets overlapped), per GPU
            spp_seg = st["n_segments"] * world / n_samples; spp_sh = st["n_shadow_rays"] * world / n_samples
            b_sample = B_RAYGEN + B_FILM + spp_seg * (B_HIT + B_SHADE + B_COMPACT) + spp_sh * B_SHADOW
            roof["bytes_per_sample"] = b_sample
            roof["step_gbs_per_gpu"] = value * 1e6 * b_sample / 1e9 / world
            roof["step_frac"] = roof["step_gbs_per_gpu"] / roof["peak"]
            line["roofline"] = roof; line["stages"] = stages
            if bvh:
                line["bvh"] = bvh
        if a.cpu_baseline and world == 1 and a.workload == "cbox":
            line["cpu_baseline"] = cpu_sample(a, 48)      # ~12 s on 128 host threads
        if a.parity and world == 1 and a.workload == "cbox":
            line["parity"] = parity_vs_reference(a, scene, integ)
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def scene_lane_slots(a, st):
    return a.lane_slots or 148 * 3072


def ctypes_sizeof_params():
    import ctypes
    from mitsuba2_b200 import abi
    return ctypes.sizeof(abi.B2RenderParams)


def parity_vs_reference(a, scene, integ):
    """--parity: render the SAME configuration with the unmodified reference CLI on this box's host cores and compare the
    developed images (rel-L2 of the fp32 RGB image, HDRFilm::bitmap semantics; flip-robust statistics as in DESIGN.md 5)."""
    from tests import golden_cases as G
    from tests.ref_cli import render_cbox
    ref, secs = render_cbox(a.res, a.spp, a.spp_per_pass, a.max_depth, usable_cores())
    if ref is None:
        return {"unavailable": "oracle/_ref/dist/mitsuba cannot run here"}
    film = integ.render(scene, a.spp)
    s = G.parity_stats(scene.develop(film), ref)
    s["reference_seconds"] = secs
    return s


_REAL_STDOUT = None


def emit(text):
    """The ONE stdout line of the contract."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(text, flush=True)


def main():
    # everything else that libraries write to fd 1 (e.g. NCCL's version banner) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        main_reference(a, rank)
    else:
        main_b2pt(a, rank, world, local)


if __name__ == "__main__":
    main()
