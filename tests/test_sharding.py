"""CPU: the multi-GPU contract - chains shard over ranks with Philox
subsequence = GLOBAL chain id (RcConfig.chain_offset), so an ensemble does not depend on how many
ranks ran it, and the one collective is the all-reduce of the ensemble histograms.  The oracle
stands in for the engine (no GPU here; the engine is bit-exact with it), gloo for NCCL."""

import json
import os
import subprocess
import sys
import textwrap
from dataclasses import replace

import numpy as np

from gerrychain_b200 import workloads
from gerrychain_b200.distributed import shard_chains

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_with_chain_offset_reproduce_the_single_run():
    from oracle import oracle

    csr, plan, cfg = workloads.config1(n_chains=12, seed=77)
    whole = oracle.OracleChains(csr, cfg, plan).run(40, n_threads=4)
    parts = []
    for r in range(3):
        off, n = shard_chains(12, r, 3)
        shard = oracle.OracleChains(csr, replace(cfg, n_chains=n, chain_offset=off), plan).run(40, n_threads=2)
        parts.append(shard)
    assert np.array_equal(np.concatenate([p.assignments() for p in parts]), whole.assignments())
    assert np.array_equal(np.concatenate([p.cut_count for p in parts]), whole.cut_count)
    assert np.array_equal(np.concatenate([p.tally for p in parts]), whole.tally)
    # and the chains of one run do differ from each other (distinct subsequences)
    assert len({a.tobytes() for a in whole.assignments()}) == 12


def test_two_ranks_over_gloo_give_the_single_process_ensemble(tmp_path):
    from oracle import oracle

    script = tmp_path / "w.py"
    out_json = tmp_path / "ens.json"
    script.write_text(textwrap.dedent(f"""
        import json, os, sys
        from dataclasses import replace
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        from gerrychain_b200 import workloads
        from gerrychain_b200.distributed import all_reduce_histograms, gather_chain_scalars, shard_chains
        from oracle import oracle
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        off, n = shard_chains(16, r, w)
        csr, plan, cfg = workloads.config1(n_chains=n, seed=5, chain_offset=off)
        ora = oracle.OracleChains(csr, cfg, plan).run(30, n_threads=2)
        cut = torch.from_numpy(np.bincount(ora.cut_count, minlength=256).astype(np.int64))
        ideal = cfg.pop_target
        worst = np.abs(ora.tally - ideal).max(axis=1) / ideal
        dev = torch.from_numpy(np.bincount(np.minimum((worst / 0.001).astype(np.int64), 63), minlength=64).astype(np.int64))
        all_reduce_histograms(cut, dev)          # the checkpoint collective (NCCL on GPUs)
        per_chain = gather_chain_scalars(torch.from_numpy(ora.cut_count.astype(np.int32)))
        if r == 0:
            json.dump(dict(cut=cut.tolist(), dev=dev.tolist(), per_chain=per_chain.tolist()), open({str(out_json)!r}, "w"))
        dist.destroy_process_group()
        print("rank%d_ok" % r, flush=True)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29641", str(script)],
        capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "rank0_ok" in out.stdout and "rank1_ok" in out.stdout
    got = json.load(open(out_json))
    csr, plan, cfg = workloads.config1(n_chains=16, seed=5)
    one = oracle.OracleChains(csr, cfg, plan).run(30, n_threads=4)
    assert got["per_chain"] == one.cut_count.tolist()
    assert got["cut"] == np.bincount(one.cut_count, minlength=256).tolist()
    worst = np.abs(one.tally - cfg.pop_target).max(axis=1) / cfg.pop_target
    assert got["dev"] == np.bincount(np.minimum((worst / 0.001).astype(np.int64), 63), minlength=64).tolist()
    assert sum(got["cut"]) == 16 and sum(got["dev"]) == 16
