"""World-size-2 test of the multi-rank host logic on CPU (gloo): chains shard with no data-path collective, the
final all-gather of packed minima gives every rank the same global winner, independent of the rank count."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(
    """
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, %(root)r)
    from transrot_b200 import chain_range, gather_minima, global_minimum
    from transrot_b200.engine import MINIMUM_DTYPE
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total = 37                                      # ragged on purpose
    lo, hi = chain_range(rank, world, n_total)
    rng = np.random.default_rng(123)
    energy_all = rng.normal(-100.0, 5.0, size=n_total)      # what each chain "finds" depends on the GLOBAL id only
    local = np.zeros(hi - lo, dtype=MINIMUM_DTYPE)
    local["energy"] = energy_all[lo:hi]
    local["chain"] = np.arange(lo, hi)
    local["tooth"] = np.arange(lo, hi) %% 5
    counts = [chain_range(r, world, n_total)[1] - chain_range(r, world, n_total)[0] for r in range(world)]
    t = torch.from_numpy(local.view(np.uint8).copy())
    allmin = gather_minima(t, world, counts=counts)
    assert len(allmin) == n_total and list(allmin["chain"]) == list(range(n_total))
    assert np.array_equal(allmin["energy"], energy_all)
    e, c, tooth = global_minimum(allmin)
    assert c == int(np.argmin(energy_all)) and tooth == c %% 5
    print(f"rank {rank} ok winner {c} {e:.6f}")
    dist.destroy_process_group()
    """
)


def test_two_rank_minima_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    import re

    hits = re.findall(r"rank (\d) ok winner (\d+) (-?\d+\.\d{6})", out.stdout)       # the two ranks may share a line
    assert sorted(h[0] for h in hits) == ["0", "1"] and len({h[1:] for h in hits}) == 1, out.stdout
