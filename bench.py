#!/usr/bin/env python
"""bench.py — retrieval queries/s of the motion-retrieval hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c0|c0loop|c1|c2|c3q1|c3q4096|c4]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)
    python bench.py --impl reference ...      # the reference's CPU call pattern (oracle port)

A step = one pass of the hot path (scan -> top-k -> fp32 re-score -> filter [-> gather]) over
one batch of synthetic queries. Default workload c1 = BASELINE.json configs[1]: 1 M-entry
768-d table, ONE query per step, k = 12 (= ref_video_num + 3), `video != own` post-filter.
With N > 1 the same table is row-sharded over the ranks (strong scaling) and a step also
contains the all-gather of per-shard candidates and the merge. Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

DIM, TOPK, K_REF, L_TOK, C_FEAT = 768, 12, 9, 25, 1024
WORKLOADS = {
    # name: (rows, queries per step, data kind, description)
    "c1": (1_000_000, 1, "clustered", "1M-entry DB, single-query top-12 (HBM-streaming scan of the bf16 shadow rows, "
                                      "fp32 re-rank from the master rows)"),
    "c2": (1_000_000, 4096, "clustered", "1M-entry DB, 4096-query batch top-12 (tcgen05 scan, fused epilogue top-k)"),
    "c3q1": (10_000_000, 1, "clustered", "10M-entry DB row-sharded, single query"),
    "c3q4096": (10_000_000, 4096, "clustered", "10M-entry DB row-sharded, 4096-query batch"),
    "c4": (1_000_000, 16, "clustered", "1M-entry DB, 16-query batch + gather into CAMA context [16,250,1024] bf16 from a "
                                       "row-aligned 1M x 25 x 1024 bf16 feature table"),
    "c0": (100_000, 64, "clustered", "100k-entry DB, 64-query batch top-12 (BASELINE config 0, the reference's CPU-runnable case; "
                                     "one padded tensor tile)"),
    "c0loop": (100_000, 1, "clustered", "100k-entry DB, one query per call (BASELINE config 0 in the reference's call pattern)"),
    "c1q4": (1_000_000, 4, "clustered", "1M-entry DB, 4 queries per step (one padded tensor tile at the HBM rate)"),
    "c1s": (125_000, 1, "clustered", "125k-entry DB, single query (the per-GPU shard of c1 at 8 GPUs; tuning aid)"),
    "c1f": (1_000_000, 1, "clustered", "1M-entry DB, single query, fp32 master rows streamed (4 B/elt, ranking exact in fp32)"),
}
PATHS = {"c1f": "stream_f32"}
POOL = 16  # distinct query batches cycled through the steps


def config_of(workload: str, world: int) -> dict:
    """The `config` object of the JSON line — built by ONE function for both arms, so the driver's
    same-config check compares like with like."""
    n_rows, nq, kind, desc = WORKLOADS[workload]
    return {"workload": f"{workload}: {desc}", "db_rows": n_rows, "dim": DIM, "queries_per_step": nq, "top_k": TOPK,
            "filter": "post-filter video != own", "data_kind": kind, "sharding": f"rows/{world}",
            "l2_flush": (f"none needed: {n_rows // world * DIM * 2 / 1e6:.0f} MB of bf16 rows streamed per GPU and step "
                         "(126 MB L2), 16 query batches cycled"), "query_pool": POOL}


def ncu_traffic(workload: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (or None)."""
    f = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(f.read_text()).get(workload, {}).get("traffic_bytes_per_launch")
    except Exception:
        return None


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        p = json.loads(f.read_text())
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.file = gpu_index, None, None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [r.split(",") for r in Path(self.file.name).read_text().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.file.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "reasons": sorted(reasons)}


# ================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist

    import motionrag_b200 as m
    from motionrag_b200 import synthetic

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pk = peaks()
    stores = {}
    # cross-GPU merge: fused into the last search kernel over peer-mapped memory (default) or
    # one NCCL all-gather + merge kernel (--exchange nccl)
    xchg = None
    if world > 1 and args.exchange == "peer":
        # peer mapping needs CUDA IPC + P2P between the ranks' GPUs; if any rank cannot set it up
        # every rank agrees to use the NCCL all-gather transport instead (same results)
        try:
            xchg = m.PeerExchange(rank, world, dev, nq_cap=4096, k_cap=32)
            ok = 1.0
        except Exception as e:   # noqa: BLE001
            print(f"[rank {rank}] peer exchange unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
            ok = 0.0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if float(flag.item()) < 1.0:
            xchg = None

    def get_store(n_rows, kind):
        key = (n_rows, kind)
        if key not in stores:
            for old in list(stores):          # one table resident at a time
                stores.pop(old)[0].close()
            feats.clear()
            torch.cuda.empty_cache()
            rps = -(-n_rows // world)
            rps = -(-rps // synthetic.CHUNK_ROWS) * synthetic.CHUNK_ROWS
            lo, hi = min(n_rows, rank * rps), min(n_rows, (rank + 1) * rps)
            st = m.EmbeddingStore(DIM, max(hi - lo, 1), dev)
            synthetic.fill_store(st, hi - lo, kind, seed=0, first_row=lo)
            st.set_groups(synthetic.groups(hi - lo, lo, dev))
            stores[key] = (st, m.ShardedRetriever(st, rank, world, rps, exchange=xchg), rps, lo, hi)
        return stores[key]

    feats = {}

    def get_features(n_feat, rows_per):
        if n_feat not in feats:
            feats.clear()
            torch.cuda.empty_cache()
            lo_f = min(n_feat, rank * rows_per)
            have = max(min(n_feat, lo_f + rows_per) - lo_f, 0)
            block = m.alloc_feature_block(max(rows_per, 1), L_TOK, C_FEAT, torch.bfloat16, dev)
            step_rows = 8192
            g = torch.Generator(device=dev).manual_seed(2 + rank)
            for s0 in range(0, have, step_rows):     # N(0,1) tokens written in place (no fp32 staging of 100 GB)
                block[s0:min(have, s0 + step_rows)].normal_(generator=g)
            torch.cuda.synchronize()
            ft = m.FeatureTable(block, rows_per_shard=rows_per, shard_rank=rank, n_shards=world, n_rows=n_feat)
            feats[n_feat] = m.open_peer_tables(ft) if world > 1 else ft
        return feats[n_feat]

    def make_queries(st, nq, seed):
        """POOL batches of nq un-normalised queries near rows of rank 0's shard, same on all ranks."""
        g = torch.Generator(device=dev).manual_seed(seed)
        src = torch.randint(0, len(st), (POOL, nq), generator=g, device=dev)
        q = synthetic.queries_from_rows(st.rows_f32()[src.flatten()], seed=seed + 1).view(POOL, nq, DIM)
        ex = (src // 3).to(torch.int32)
        if world > 1:
            dist.broadcast(q, 0)
            dist.broadcast(ex, 0)
        return q.contiguous(), ex.contiguous()

    def recheck_count(st, retr, q, ex, spath, nq):
        """How many queries of ONE pass over the pooled batches end up on the fp32 master rows when they go
        through the certified host-facing API (RAGDatabase: margin test, deeper lists, then fp32)."""
        db = m.RAGDatabase.from_store(st, {}, retriever=retr if world > 1 else None, path=spath)
        qn, exn = q.cpu().numpy(), ex.cpu().numpy()
        for j in range(POOL if nq <= 64 else 2):
            db.search_arrays(qn[j], TOPK, exclude_group=exn[j])
        return {"fp32": int(db.fp32_rechecks), "deeper_lists": int(db.deep_rechecks),
                "queries": int((POOL if nq <= 64 else 2) * nq)}

    def parity_probe(name, st, retr, q, ex, spath, lo, hi, max_queries=64):
        """Correctness evidence inside the bench, at every N: up to 64 pooled queries through the very call the
        timed loop makes, against a float64 brute force over ALL shards (per-rank fp64 GEMM on the fp32 master
        rows, candidates all-gathered, merged, post-filtered). Rule of BASELINE.md §5: a differing index is a
        near-tie when the exact distances differ by < 1e-3 relative, anything else a mismatch (must be 0)."""
        nq = q.shape[1]
        per = min(nq, max_queries)                                   # queries checked per batch
        take = max(1, min(POOL, max_queries // per))                 # batches run (whole, as in the timed steps)
        got_i, got_d = [], []
        for j in range(take):
            r = retr.search(q[j], TOPK, path=spath, exclude_group=ex[j], filter_mode="post")
            got_i.append(r.index[:per].clone())
            got_d.append(r.distance[:per].clone())
        got_i, got_d = torch.cat(got_i), torch.cat(got_d)
        qs = q[:take, :per].reshape(-1, DIM).contiguous()
        exs = ex[:take, :per].reshape(-1).contiguous()
        rows = st.rows_f32()
        qd = qs.double()
        qq = (qd * qd).sum(-1, keepdim=True)
        best_d = torch.full((qs.shape[0], TOPK), float("inf"), dtype=torch.float64, device=dev)
        best_i = torch.full((qs.shape[0], TOPK), -1, dtype=torch.int64, device=dev)
        for s0 in range(0, rows.shape[0], 1 << 18):
            r64 = rows[s0:s0 + (1 << 18)].double()
            d = qq + (r64 * r64).sum(-1)[None] - 2.0 * (qd @ r64.T)
            cd, ci = torch.topk(d, min(TOPK, d.shape[1]), dim=-1, largest=False)
            alld, alli = torch.cat([best_d, cd], 1), torch.cat([best_i, ci + s0 + lo], 1)
            o = torch.argsort(alld, dim=-1, stable=True)[:, :TOPK]
            best_d, best_i = alld.gather(1, o), alli.gather(1, o)
        if world > 1:
            gd = [torch.empty_like(best_d) for _ in range(world)]
            gi = [torch.empty_like(best_i) for _ in range(world)]
            dist.all_gather(gd, best_d)
            dist.all_gather(gi, best_i)
            alld, alli = torch.cat(gd, 1), torch.cat(gi, 1)
            o = torch.argsort(alld, dim=-1, stable=True)[:, :TOPK]
            best_d, best_i = alld.gather(1, o), alli.gather(1, o)
        # post-filter on the exact list: drop rows of the query's own video (group = row // 3), keep order
        keep = (best_i // 3) != exs[:, None].long()
        o = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)       # kept entries first, order preserved
        counts = keep.sum(1)
        ar = torch.arange(TOPK, device=dev)[None]
        ref_i = torch.where(ar < counts[:, None], best_i.gather(1, o), torch.full_like(best_i, -1))
        ref_d = torch.where(ar < counts[:, None], best_d.gather(1, o), torch.full_like(best_d, float("inf")))
        same_count = ((got_i >= 0).sum(1) == counts)
        valid = (got_i >= 0) & (ref_i >= 0)
        diff = (got_i != ref_i) & valid
        rel = (got_d.double() - ref_d).abs() / ref_d.abs().clamp_min(1e-3)
        near = diff & (rel <= 1e-3)
        # distances of positions that agree must be the exact fp32 values (1e-3 is the contract; fp32 gives ~1e-6)
        err = torch.where(valid & ~diff, rel, torch.zeros_like(rel)).max()
        return {"queries": int(qs.shape[0]), "checked": int(valid.sum()), "mismatches": int((diff & ~near).sum()) + int((~same_count).sum()),
                "near_ties": int(near.sum()), "max_rel_distance_err": float(err), "oracle": "float64 brute force over all shards, "
                "all-gathered; post-filter applied to the exact top-12"}

    def measure(name, steps, warmup, with_e2e=True, sample_clocks=False):
        n_rows, nq, kind, desc = WORKLOADS[name]
        st, retr, rps, lo, hi = get_store(n_rows, kind)
        q, ex = make_queries(st, nq, seed=100 + nq)
        gather = name == "c4"
        spath = PATHS.get(name, "auto")
        ctx = None
        if gather:
            # the feature table is ROW-ALIGNED with the embedding table (row i = motion tokens of clip i,
            # 51 200 B each: 51.2 GB at 1 M rows), row-sharded like it: each rank owns its rows in an
            # IPC-exportable block and the gather kernel reads peers' rows over NVLink through mapped pointers
            n_feat = n_rows
            ftable = get_features(n_feat, rps)
            gg = torch.Generator(device=dev).manual_seed(4)
            sos = (torch.randn(1, L_TOK, C_FEAT, generator=gg, device=dev) / 32).bfloat16()
            un = torch.randn(L_TOK, C_FEAT, generator=gg, device=dev).bfloat16()
            cond = torch.randn(nq, (K_REF + 1) * L_TOK, C_FEAT, generator=gg, device=dev).bfloat16()
            ctx = m.MotionContext(ftable, sos, un, pe_max_length=256)

        margins = []     # exactness margin of every timed query (device tensors, judged after the timed region)
        certify = spath != "stream_f32"

        def step(i, timings=None):
            j = i % POOL
            if timings is not None and world == 1:
                r = st.search(q[j], TOPK, path=spath, exclude_group=ex[j], filter_mode="post", timings=timings)
            else:
                r = retr.search(q[j], TOPK, path=spath, exclude_group=ex[j], filter_mode="post", certify=certify)
                if certify and len(margins) < 4096:
                    margins.append(r.margin)
            if gather:
                return ctx.build(r.index[:, :K_REF].contiguous(), cond)
            return r

        for i in range(warmup):
            step(i)
        barrier()
        margins.clear()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        launches0 = m.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        launches = m.launch_count() - launches0
        barrier()
        ms_total = allmax(e0.elapsed_time(e1))
        clocks = None
        if sampler:
            # nvidia-smi samples every 100 ms; a short timed region would yield no sample at all, so
            # the identical step loop keeps running (untimed; same count on every rank) until about
            # 0.6 s of load has been observed
            extra_steps = int(min(20000, max(0, 600.0 - ms_total) / max(ms_total / steps, 1e-3)))
            for i in range(extra_steps):
                step(i)
            torch.cuda.synchronize()
            clocks = sampler.stop()
            clocks["window"] = "timed region + identical untimed continuation of the same loop to ~0.6 s"
            barrier()
        out = {"workload": name, "desc": desc, "db_rows": n_rows, "queries_per_step": nq, "steps": steps,
               "ms_per_step": ms_total / steps, "value": steps * nq / (ms_total / 1e3), "gpu_launches": launches}
        # queries of the timed region whose bf16 scan is NOT certified exact by its margin (6-sigma test): the
        # host-facing API re-runs such a query on the fp32 rows; the device-resident API hands the margin back
        if certify and margins:
            from motionrag_b200.store import PATH_NAME, margin_threshold
            used = PATH_NAME[st.plan(nq, k=TOPK, path=spath).path]
            thr = margin_threshold(used, DIM, False, st.info().max_norm_deviation)
            mg = torch.cat([t.flatten() for t in margins])
            out["uncertified"] = int((~(mg > thr)).sum())     # margin below the 6-sigma threshold in the timed region
            out["certified_queries"] = int(mg.numel())
            out["margin_min_over_threshold"] = float(mg.min() / thr)
            out["fp32_rechecks"] = recheck_count(st, retr, q, ex, spath, nq)
        else:
            out["fp32_rechecks"] = out["uncertified"] = 0
        out["parity"] = parity_probe(name, st, retr, q, ex, spath, lo, hi)
        # per-step latency distribution (device time per step, max over ranks)
        lat = []
        for i in range(min(steps, 200)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            a.record()
            step(i)
            b.record()
            b.synchronize()
            lat.append(a.elapsed_time(b))
        lat_t = torch.tensor(lat, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)
        lat = sorted(lat_t.tolist())
        out["p50_ms"] = lat[len(lat) // 2]
        out["p95_ms"] = lat[min(len(lat) - 1, int(len(lat) * 0.95))]
        # roofline of the dominant (scan) kernel: event-bracketed launches on this rank's shard
        tm = []
        for i in range(min(steps, 50)):
            st.search(q[i % POOL], TOPK, path=spath, exclude_group=ex[i % POOL], filter_mode="post", timings=tm)
        scan_ms = allmax(statistics.mean(t[0] for t in tm))
        plan = st.plan(nq, k=TOPK, filter_mode="post", path=spath)
        if plan.path == 3 and plan.m_tiles == 1:
            # one (padded) 128-query tile: the tensor kernel streams the bf16 table once -> HBM-bound
            nbytes = plan.scan_bytes
            ach = nbytes / (scan_ms / 1e3) / 1e9
            out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": ach / pk["hbm_gbs"], "traffic": None, "kernel": "k2_batch_kernel",
                               "kernel_ms": scan_ms, "kernel_share_of_step": scan_ms / (ms_total / steps),
                               "peak_source": pk["source"], "bytes_per_launch": nbytes,
                               "plan": {"grid": plan.grid, "m_tiles": plan.m_tiles, "n_tiles": plan.n_tiles,
                                        "runs": plan.chunks}}
        elif plan.path == 3:
            ach = plan.scan_flops / (scan_ms / 1e3) / 1e12
            peak = pk["bf16_tflops_sustained"] if steps * scan_ms > 2000 else pk["bf16_tflops"]
            out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                               "frac": ach / peak, "traffic": ncu_traffic(name) if world == 1 else None,
                               "kernel": "k2_batch2_kernel" if nq > 128 else "k2_batch_kernel",
                               "kernel_ms": scan_ms, "kernel_share_of_step": scan_ms / (ms_total / steps),
                               "peak_source": pk["source"],
                               "plan": {"grid": plan.grid, "m_tiles": plan.m_tiles, "n_tiles": plan.n_tiles,
                                        "chunks": plan.chunks}}
        else:
            # Single-query searches are ONE kernel launch per step (scan + fused tail), and consecutive launches
            # pipeline (programmatic dependent launch: the scan of search i+1 runs under the tail of search i). The
            # kernel's average launch duration over the timed region is therefore the step time itself; the
            # event-bracketed duration of one ISOLATED launch (kernel_ms_isolated, nothing to overlap with) is kept
            # beside it. The peak is a COPY bandwidth (read + write); a read-only stream can exceed it.
            one_launch_steps = launches == steps and not gather
            in_loop_ms = ms_total / steps if one_launch_steps else scan_ms
            ach = plan.scan_bytes / (in_loop_ms / 1e3) / 1e9
            ach_iso = plan.scan_bytes / (scan_ms / 1e3) / 1e9
            out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": ach / pk["hbm_gbs"], "traffic": ncu_traffic(name) if world == 1 else None,
                               "traffic_source": "static: ncu capture committed under profiles/ (profiles/traffic.json)",
                               "kernel": "k1_stream_kernel<float>" if plan.path == 1 else "k1_stream_kernel<bf16>",
                               "kernel_ms": in_loop_ms, "kernel_ms_isolated": scan_ms, "frac_isolated": ach_iso / pk["hbm_gbs"],
                               "kernel_share_of_step": in_loop_ms / (ms_total / steps),
                               "timing": ("average launch duration over the timed region (launches pipeline)" if one_launch_steps
                                          else "event-bracketed single launches"),
                               "peak_source": pk["source"], "peak_nominal_hbm3e_gbs": 8000.0,
                               "bytes_per_launch": plan.scan_bytes,
                               "plan": {"grid": plan.grid, "cands_per_query": plan.cands_per_query}}
        out["total_ms_single_rank_call"] = statistics.mean(t[1] for t in tm)
        if clocks:
            out["clocks"] = clocks

        # end to end: HOST query buffers in, HOST results out, every step
        if with_e2e:
            q_host = q.cpu().pin_memory()
            ex_host = ex.cpu().pin_memory()
            import numpy as np
            db = None
            if not gather and nq == 1 and spath == "auto":
                # the reference-facing call itself, at every N: RAGDatabase.text_search(ndarray) -> list[dict]
                # (row-sharded tables: every rank makes the same call; the peer exchange runs inside the
                # captured graph of mrag_search_sharded_host)
                cols = {"video": np.array([f"video_{j // 3:07d}.mp4" for j in range(n_rows)]),
                        "start_sec": np.zeros(n_rows), "end_sec": np.ones(n_rows) * 2}
                db = m.RAGDatabase.from_store(st, cols, retriever=retr if world > 1 else None)
                qn = q_host.numpy()
                wh = [f'video != "video_{int(ex_host[j, 0]):07d}.mp4"' for j in range(POOL)]

                def e2e_step(i):
                    j = i % POOL
                    return db.text_search(qn[j, 0], top_k=TOPK, where=wh[j], select=["video", "start_sec", "end_sec"])
                api = "RAGDatabase.text_search(ndarray[768]) -> list[dict]"
                d2h = TOPK * 16 + 4
            elif not gather:
                # batches: the certified array-level call (host arrays in / out; queries whose margin fails are
                # re-issued with 32-entry lists, then from the fp32 rows — inside the timed step)
                qn, exn = q_host.numpy(), ex_host.numpy()
                db = m.RAGDatabase.from_store(st, {}, retriever=retr if world > 1 else None, path=spath)

                def e2e_step(i):
                    j = i % POOL
                    return db.search_arrays(qn[j], TOPK, exclude_group=exn[j])
                api = "RAGDatabase.search_arrays(ndarray[nq,768]) -> ndarrays (distance, index), certified"
                d2h = nq * (TOPK * 16 + 4)
            else:
                def e2e_step(i):
                    j = i % POOL
                    qd = q_host[j].to(dev, non_blocking=True)
                    exd = ex_host[j].to(dev, non_blocking=True)
                    r = retr.search(qd, TOPK, path=spath, exclude_group=exd, filter_mode="post")
                    x = ctx.build(r.index[:, :K_REF].contiguous(), cond)
                    return x.float().sum().item(), r.index.cpu()
                api = "ShardedRetriever.search(pinned host queries) -> host index + gather_context checksum"
                d2h = nq * TOPK * 8 + 4
            for i in range(max(3, warmup // 2)):
                e2e_step(i)
            barrier()
            t0 = time.perf_counter()
            lat = []
            for i in range(steps):
                t1 = time.perf_counter()
                e2e_step(i)
                lat.append(time.perf_counter() - t1)
            torch.cuda.synchronize()
            dt = allmax(time.perf_counter() - t0)
            out["e2e"] = {"value": steps * nq / dt, "unit": "queries/s", "ms_per_step": dt / steps * 1e3,
                          "p50_ms": statistics.median(lat) * 1e3, "max_ms": max(lat) * 1e3,
                          "h2d_bytes_per_step": nq * (DIM * 4 + 4) + 4, "d2h_bytes_per_step": d2h, "api": api}
            if db is not None:   # queries whose bf16 scan was not certified: re-issued with deeper lists / in fp32
                out["e2e"]["fp32_rechecks"] = int(db.fp32_rechecks)
                out["e2e"]["deep_rechecks"] = int(db.deep_rechecks)
        return out

    def measure_gather(steps=20, warmup=3):
        """K4 alone on the ROW-ALIGNED 1 M x 25 x 1024 bf16 feature table (51.2 GB), at the batch sizes the
        reference's configs name (b = 1: CogVideoX inference, b = 16: motion-transformer eval) and at bulk size
        (b = 4096): feature rows -> [b,250,1024] bf16 context with +pe and +cond fused.
        Bytes per sample: 9 rows x 51 200 B read + 512 000 B cond read + 512 000 B written."""
        for old in list(stores):
            stores.pop(old)[0].close()
        torch.cuda.empty_cache()
        n_feat = 1_000_000
        ft = get_features(n_feat, -(-n_feat // world))
        gg = torch.Generator(device=dev).manual_seed(4)
        sos = (torch.randn(1, L_TOK, C_FEAT, generator=gg, device=dev) / 32).bfloat16()
        un = torch.randn(L_TOK, C_FEAT, generator=gg, device=dev).bfloat16()
        ctx = m.MotionContext(ft, sos, un, pe_max_length=256)
        res = {"workload": "K4 gather_context from a row-aligned 1M-row feature table (51.2 GB bf16), K=9, +pe +cond fused",
               "kernel": "k4_gather_kernel<bf16>"}
        row = L_TOK * C_FEAT * 2
        for b in (1, 16, 4096):
            cond = torch.randn(b, (K_REF + 1) * L_TOK, C_FEAT, generator=gg, device=dev).bfloat16()
            idx = torch.randint(0, n_feat, (64, b, K_REF), generator=gg, device=dev)   # 64 index sets: rows come from HBM
            idx[:, ::7, 3] = -1
            out = torch.empty(b, (K_REF + 1) * L_TOK, C_FEAT, dtype=torch.bfloat16, device=dev)
            n = steps if b == 4096 else 200
            for i in range(warmup):
                ctx.build(idx[i % 64], cond, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                ctx.build(idx[i % 64], cond, out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            nbytes = b * (K_REF * row + 2 * (K_REF + 1) * row)
            ach = nbytes / (ms / 1e3) / 1e9
            res[f"b{b}"] = {"ms_per_step": ms, "samples_per_s": b / (ms / 1e3),
                            "roofline": {"bound": "hbm" if b >= 256 else "latency (one short wave: %d CTAs)" % (b * 10 * 4),
                                         "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                                         "bytes_per_launch": nbytes}}
            # the kernel without the host: 16 launches (16 different index sets) captured once, replayed as a graph.
            # At b = 1 / 16 the eager figure above is the Python + launch cost of one call, not the kernel.
            try:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    ctx.build(idx[0], cond, out)
                torch.cuda.current_stream(dev).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    for i in range(16):
                        ctx.build(idx[i], cond, out)
                reps = 3 if b == 4096 else 50
                for _ in range(2):
                    graph.replay()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(reps):
                    graph.replay()
                e1.record()
                torch.cuda.synchronize()
                ms_k = e0.elapsed_time(e1) / (reps * 16)
                res[f"b{b}"]["kernel_ms_graph_replay"] = ms_k
                res[f"b{b}"]["roofline"]["frac_graph_replay"] = nbytes / (ms_k / 1e3) / 1e9 / pk["hbm_gbs"]
                del graph
            except Exception as e:   # noqa: BLE001
                res[f"b{b}"]["graph_replay_error"] = f"{type(e).__name__}: {e}"[:200]
        feats.clear()
        torch.cuda.empty_cache()
        return res

    def measure_bulk(n_anno=16384):
        """The reference's actual bulk use (prepare_annotations, src/data/datamodule.py:231-265):
        one `ref_videos` record list per annotation, host embeddings in, Python dicts out —
        through RAGDatabase.retrieve_for_annotations (batched scans of 4096 queries)."""
        import numpy as np
        n_rows = 1_000_000
        st, retr, rps, lo, hi = get_store(n_rows, "clustered")
        cols = {"video": np.array([f"video_{j // 3:07d}.mp4" for j in range(n_rows)]),
                "start_sec": np.zeros(n_rows), "end_sec": np.ones(n_rows) * 2}
        db = m.RAGDatabase.from_store(st, cols)
        g = torch.Generator(device=dev).manual_seed(11)
        src = torch.randint(0, n_rows, (n_anno,), generator=g, device=dev)
        emb = synthetic.queries_from_rows(st.rows_f32()[src], seed=12).cpu().numpy()
        vids = cols["video"][src.cpu().numpy()]
        annos = [{"video": str(v), "text_embedding": e} for v, e in zip(vids, emb)]
        db.retrieve_for_annotations([dict(a) for a in annos[:4096]], K_REF)          # warm-up
        t0 = time.perf_counter()
        out = db.retrieve_for_annotations(annos, K_REF)
        dt = time.perf_counter() - t0
        n_ref = sum(len(a["ref_videos"]) for a in out)
        return {"workload": f"prepare_annotations rag_text branch: {n_anno} annotations x 1M-entry DB, k = K+3 = 12, "
                            "`video != own`, records attached as anno['ref_videos']",
                "value": n_anno / dt, "unit": "annotations/s", "seconds": dt, "records": n_ref,
                "api": "RAGDatabase.retrieve_for_annotations(list[dict]) -> list[dict]"}

    def measure_cama(steps=50, warmup=5):
        """the CAMA causal transformer forward (4 post-norm layers, d 1024, 16 heads,
        ff 4096, block-causal 10 x 25 mask) on libmrag kernels K5/K6/K7, replayed as one CUDA graph; and the
        whole device-side chain retrieval -> gather -> forward for one query. Random-init weights."""
        import torch.nn as nn
        for old in list(stores):
            stores.pop(old)[0].close()
        torch.cuda.empty_cache()
        torch.manual_seed(0)
        layer = nn.TransformerEncoderLayer(C_FEAT, 16, 4096, 0.0, "gelu", batch_first=True, norm_first=False)
        enc = nn.TransformerEncoder(layer, 4, enable_nested_tensor=False)
        cama = m.CamaTransformer(enc, groups=K_REF + 1, group_tokens=L_TOK, max_batch=16, device=dev.index or 0)
        T = (K_REF + 1) * L_TOK
        res = {"workload": "CAMA forward: 4 layers, d_model 1024, 16 heads, d_ff 4096, 250 tokens, bf16, CUDA-graph replay",
               "kernels": "k5_linear_kernel / k5_linear_pair_kernel (tcgen05), k6_attention_kernel, k7_add_layernorm_kernel; 28 launches/forward"}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timed(fn, n):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1e3

        def gemm_flop(rows):      # the four GEMMs of one layer over `rows` token rows
            return 2.0 * rows * C_FEAT * (3 * C_FEAT + C_FEAT + 2 * 4096)

        def attn_flop(b, groups):  # block-causal attention: group gi sees (gi+1)*L keys; QK^T and PV
            return sum(b * 16 * 2 * (2.0 * L_TOK * (gi + 1) * L_TOK * 64) for gi in groups)

        enc_bf16 = None
        xs_all = {b: torch.randn(b, T, C_FEAT, device=dev).bfloat16() for b in (1, 16)}   # before any graph capture
        for b in (1, 16):
            cama.input_view(b).copy_(xs_all[b])
            us = timed(lambda: cama.forward(b=b), steps)
            M = b * T
            flop = 4 * (gemm_flop(M) + attn_flop(b, range(K_REF + 1)))
            tf = flop / (us * 1e-6) / 1e12
            res[f"b{b}"] = {"us_per_forward": us, "samples_per_s": b / (us * 1e-6), "tflops": tf,
                            "frac_of_tensor_peak": tf / pk["bf16_tflops"],
                            "roofline": {"bound": "tensor" if b >= 16 else "latency (28 dependent launches of < 1 wave each)",
                                         "achieved": tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops"],
                                         "flop_per_forward": flop, "peak_source": pk["source"]}}
            # predict(): what ActionTransformer.predict needs (module.py:326 keeps the last group only) — the last
            # layer runs attention / out-proj / FFN / LayerNorms on that group's 25 rows per sample; executed flops
            us_p = timed(lambda: cama.predict(b=b), steps)
            flop_p = (3 * (gemm_flop(M) + attn_flop(b, range(K_REF + 1))) + 2.0 * M * C_FEAT * 3 * C_FEAT
                      + attn_flop(b, [K_REF]) + 2.0 * b * L_TOK * C_FEAT * (C_FEAT + 2 * 4096))
            res[f"b{b}"]["predict"] = {"us": us_p, "samples_per_s": b / (us_p * 1e-6), "executed_tflops": flop_p / (us_p * 1e-6) / 1e12,
                                       "frac_of_tensor_peak": flop_p / (us_p * 1e-6) / 1e12 / pk["bf16_tflops"]}
            # torch baseline: the reference's own module (torch.nn.TransformerEncoder, bf16) under the same mask,
            # eager and replayed as ONE CUDA graph (so launch overhead is taken out of the comparison)
            try:
                import copy
                if enc_bf16 is None:
                    enc_bf16 = copy.deepcopy(enc).to(dev).bfloat16().eval()
                mask = m.context.block_causal_mask(K_REF + 1, L_TOK, dev)
                xs = xs_all[b]
                with torch.no_grad():
                    # is_causal=False: without it the module compares the mask with a triangular one on every call
                    # (a device->host sync, illegal inside a capture); the mask itself is still applied
                    res[f"b{b}"]["torch_bf16_eager_us"] = timed(lambda: enc_bf16(xs, mask, is_causal=False), max(10, steps // 2))
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        for _ in range(3):
                            enc_bf16(xs, mask, is_causal=False)
                    torch.cuda.current_stream(dev).wait_stream(side)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        ys = enc_bf16(xs, mask, is_causal=False)
                    res[f"b{b}"]["torch_bf16_cuda_graph_us"] = timed(graph.replay, steps)
                    del graph, ys
            except Exception as e:   # noqa: BLE001
                res[f"b{b}"]["torch_baseline_error"] = f"{type(e).__name__}: {e}"[:200]
        del enc_bf16
        # one query end to end on the device: scan + select + gather into the transformer's input + forward
        st, retr, rps, lo, hi = get_store(1_000_000, "clustered")
        q, ex = make_queries(st, 1, 21)                   # [POOL, 1, DIM], own-group ids [POOL, 1]
        ftab = get_features(1_000_000, rps)               # row-aligned with the embedding table
        gg = torch.Generator(device=dev).manual_seed(4)
        sos = (torch.randn(1, L_TOK, C_FEAT, generator=gg, device=dev) / 32).bfloat16()
        un = torch.randn(L_TOK, C_FEAT, generator=gg, device=dev).bfloat16()
        cond = torch.randn(1, T, C_FEAT, generator=gg, device=dev).bfloat16()
        ctx = m.MotionContext(ftab, sos, un, pe_max_length=256)

        def chain(i):
            r = st.search(q[i % POOL], TOPK, exclude_group=ex[i % POOL])
            ctx.build(r.index[:, :K_REF].contiguous(), cond, out=cama.input_view(1))
            return cama.predict(b=1)
        for i in range(warmup):
            chain(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            chain(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / steps * 1e3
        res["query_to_prediction"] = {"us": us, "what": "1 query: K1 scan of 1 M rows + K3 + K4 gather into the transformer "
                                      "input + 4-layer forward, all on one stream, inputs resident"}
        cama.close()
        return res

    main = measure(args.workload, args.steps, args.warmup, with_e2e=True, sample_clocks=True)
    extra = {}
    if not args.no_extras:
        only = set(filter(None, args.extras.split(",")))
        todo = [w for w in ("c0", "c0loop", "c1", "c1q4", "c1f", "c2", "c4", "c3q1", "c3q4096") if w != args.workload]
        for w in todo:
            if only and w not in only:
                continue
            try:
                steps = 200 if WORKLOADS[w][1] == 1 else (30 if WORKLOADS[w][1] <= 64 else 8)
                extra[w] = measure(w, steps, 3, with_e2e=(w in ("c0", "c0loop", "c2", "c4")))
            except Exception as e:  # an extra must never take the headline line down
                extra[w] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1:
            for key, fn in (("bulk_annotations", measure_bulk), ("k4_gather", measure_gather),
                            ("cama_forward", measure_cama)):
                if only and key not in only:
                    continue
                try:
                    extra[key] = fn()
                except Exception as e:
                    extra[key] = {"error": f"{type(e).__name__}: {e}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload, budget_s=12.0)

    if rank == 0:
        line = {"metric": "retrieval queries/sec", "value": main["value"], "unit": "queries/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if args.workload == "c1f" else "bf16 scan + f32 re-rank",
                "data": "synthetic (seeded clustered unit vectors, un-normalised queries; random-init features)",
                "config": config_of(args.workload, world),
                "exchange": ("none" if world == 1 else
                             ("peer-memory exchange fused into the last search kernel (NVLink stores + flags)" if xchg is not None
                              else "NCCL all_gather_into_tensor + merge kernel")),
                "p50_latency_ms": main["p50_ms"], "p95_latency_ms": main["p95_ms"],
                "e2e": main.get("e2e"), "gpu_launches": main["gpu_launches"], "roofline": main["roofline"],
                "parity": main["parity"], "fp32_rechecks": main["fp32_rechecks"],
                "cpu_baseline": cpu, "clocks": main.get("clocks"),
                # BASELINE config 3 (10 M rows row-sharded over the ranks) next to the headline workload
                "scale_10m": {w: {k2: extra[w].get(k2) for k2 in ("value", "ms_per_step", "p50_ms", "roofline", "parity",
                                                                  "fp32_rechecks", "error") if k2 in extra[w]}
                              for w in ("c3q1", "c3q4096") if w in extra},
                "c2": ({k2: extra["c2"].get(k2) for k2 in ("value", "ms_per_step", "p50_ms", "roofline", "parity", "fp32_rechecks",
                                                            "e2e", "error") if k2 in extra["c2"]} if "c2" in extra else None),
                "extra": extra}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ================================================================================================
def cpu_baseline(workload: str, budget_s: float = 12.0, steps: int | None = None, warmup: int = 3):
    """The reference's call pattern (one fp32 flat scan per query) on this host's cores — ALL of them, whatever
    OMP_NUM_THREADS the launcher exported — on a bounded sample of the same workload: the same clustered table
    and near-row queries the GPU arm uses (capped at 2 M rows for host RAM and generation time), as many queries
    as fit the budget. `steps` (reference arm): that many steps of the workload's batch (<= 64 queries each),
    but never less than 5 s or 200 queries of timed work."""
    import torch

    from motionrag_b200 import synthetic
    from oracle import cpu_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_rows, nq, kind, desc = WORKLOADS[workload]
    n = min(n_rows, 2_000_000)
    db = synthetic.database(n, DIM, kind, seed=0, device="cpu")
    db_sq = (db * db).sum(-1)
    groups = (torch.arange(n) // 3).to(torch.int32)
    g = torch.Generator().manual_seed(100)
    src = torch.randint(0, n, (4096,), generator=g)
    q = synthetic.queries_from_rows(db[src], seed=101)
    ex = groups[src]
    loop = nq == 1 or workload == "c4"
    per_call = 1 if loop else min(nq, 64)

    def call(i):
        s0 = (i * per_call) % (4096 - per_call + 1)
        if loop:   # per-query loop: exactly the reference's behaviour (src/data/datamodule.py:257-262)
            cpu_port.search_loop(db, db_sq, q[s0:s0 + 1], TOPK, groups, ex[s0:s0 + 1])
        else:
            cpu_port.search_batched(db, db_sq, q[s0:s0 + per_call], TOPK)
    for i in range(max(3, warmup)):
        call(i)
    calls, t0 = 0, time.perf_counter()
    if steps is None:
        while time.perf_counter() - t0 < budget_s and calls < 100000:
            call(calls)
            calls += 1
    else:
        want = steps * max(1, (min(nq, 64) if not loop else nq) // per_call)
        while calls < want or (time.perf_counter() - t0 < 5.0 and calls * per_call < 200):
            call(calls)
            calls += 1
            if time.perf_counter() - t0 > 120.0:
                break
    dt = time.perf_counter() - t0
    done = calls * per_call
    kind_s = ("per-query loop (reference call pattern)" if loop else
              f"batched sgemm + topk, {per_call} queries per call (best-case CPU)")
    return {"value": done / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{done} queries against a {n}-row x {DIM} fp32 {kind} table, {kind_s}, torch {torch.__version__} fp32 "
                      f"with {torch.get_num_threads()} threads, {dt:.1f} s; os.cpu_count()={os.cpu_count()}",
            "ms_per_query": dt / max(done, 1) * 1e3, "rows": n, "threads": torch.get_num_threads()}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path. LanceDB 0.14 is not
    installable offline, so this is the oracle port (oracle/cpu_port.py) with every host core.
    Rank 0 only; other ranks exit 0."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    n_rows, nq, kind, desc = WORKLOADS[args.workload]
    cpu = cpu_baseline(args.workload, steps=max(1, args.steps), warmup=max(3, args.warmup))
    per_step = nq if nq <= 64 else 64
    ms = cpu["ms_per_query"] * per_step
    line = {"impl": "reference", "metric": "retrieval queries/sec", "value": cpu["value"], "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (seeded clustered unit vectors, un-normalised queries; random-init features)",
            "config": config_of(args.workload, args.gpus),
            "note": "LanceDB 0.14 (the reference's engine) is not installable offline; this arm is the fp32 CPU "
                    "restatement of its flat scan, one scan per query like src/data/rag.py:54, on all host cores",
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=list(WORKLOADS))
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--extras", default="", help="comma-separated subset of the extra measurements (default: all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
