"""Multi-GPU parity (needs >= 2 visible GPUs; skipped otherwise): one process per GPU over NCCL.
Sharded deposit + all-reduce + replicated solve must reproduce the unsharded oracle run, and every
rank must hold bit-identical field coefficients."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from __graft_entry__ import load_package
from oracle import vm_oracle as orc
vm = load_package()
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
ctx = vm.init_distributed_context()
assert ctx.comm_info() == (rank, world)
rng = np.random.default_rng(5)
n, k, npart, dt, nt = 32, 4, 40001, 0.1, 6
a, b = 0.0, 2 * np.pi / 0.3
x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
lo, hi = vm.shard_bounds(npart, rank, world)
fld = vm.DeviceField(ctx, a, b, k, n, 0)
p = vm.DeviceParticles(ctx, hi - lo)
p.upload(x[lo:hi], v[lo:hi], w[lo:hi])
diag = fld.run(p, dt, nt, 2, 0, 1.0)
xg, vg, _ = p.download(w=False)
S = orc.periodic_stiffness(a, b, n, k, 0)
xo, vo = x.copy(), v.copy()
dref, phiref = orc.integrate_vp(xo, vo, w, dt, 1.0, nt, 2, a, b, n, k, 0, S, want_phi=True)
assert np.max(np.abs(xg - xo[lo:hi])) <= 1e-11 and np.max(np.abs(vg - vo[lo:hi])) <= 1e-11
assert np.allclose(diag[:, :3], dref, rtol=1e-10, atol=1e-13)
phi = torch.from_numpy(fld.coefficients.copy()); ref = phi.clone(); dist.broadcast(ref, src=0)
assert torch.equal(phi, ref), "field coefficients differ between ranks"
assert np.max(np.abs(phi.numpy() - phiref[-1])) <= 1e-10 * np.max(np.abs(phiref[-1]))
# v-space: sharded CLB right-hand side
vs = vm.DeviceVSpline(ctx, -10.0, 10.0, 41, 4, 1)
wv = np.full(npart, 1.0 / npart)
p.upload(v=v[lo:hi], w=wv[lo:hi])
vdot = vs.lb_rhs(p, 1.0, True)
M = orc.dirichlet_mass(-10.0, 10.0, 41, 4)
vref, _, _ = orc.lb_rhs(v, wv, -10.0, 10.0, 41, 4, M, 1.0, True)
assert np.max(np.abs(vdot - vref[lo:hi])) <= 1e-10 * np.max(np.abs(vref))
dist.barrier()
print("ok", rank)
"""


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.timeout(600)
@pytest.mark.parametrize("collective", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_run_matches_oracle(tmp_path, world, collective):
    """collective = "peer": fused deposit + NVLink peer-memory exchange + solve (one kernel per step);
    "nccl": ncclAllReduce between the deposit and the replicated solve."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, VLASOV_B200_NO_PEER="1" if collective == "nccl" else "0")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), str(script)],
                       capture_output=True, text=True, timeout=580, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world
