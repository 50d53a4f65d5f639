"""CPU-only tests: the C-ABI library loads and exports every declared symbol, fails loudly
without a device, host-side sharding logic, and the world-size-2 reduction path over gloo."""
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(vm):
    lib = vm.lib()
    header = (ROOT / "include" / "vlasov_b200.h").read_text()
    declared = set(re.findall(r"\b(vm_[A-Za-z0-9_]+)\s*\(", header))
    declared -= {"vm_status", "vm_fill_kind", "vm_deposit_mode", "vm_run_flags"}
    assert len(declared) >= 45
    for name in sorted(declared):
        assert hasattr(lib, name), f"libvlasov_b200.so does not export {name}"
    assert set(vm._lib.SIGNATURES) == declared
    assert lib.vm_abi_version() == 2


def test_no_cpu_fallback(vm):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vm.VMError) as ei:
        vm.Context(0)
    assert ei.value.code == vm._lib.VM_ERR_NO_DEVICE
    assert "no CPU path" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = ROOT / "vlasovmethods.jl_b200"
    for f in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        if f.is_file() and f.suffix in (".py", ".cu", ".cuh", ".hpp"):
            txt = f.read_text()
            assert "vm_oracle" not in txt and "oracle/" not in txt, f


def test_shard_bounds(vm):
    for n, world in [(10, 3), (100000001, 8), (5, 8), (0, 2)]:
        spans = [vm.shard_bounds(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 == b0
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        vm.shard_bounds(10, 3, 3)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from __graft_entry__ import load_package
from oracle import vm_oracle as orc
vm = load_package()
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(5)
n, k, npart = 16, 4, 20001
a, b = 0.0, 2 * np.pi / 0.3
x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
lo, hi = vm.shard_bounds(npart, rank, world)
S = orc.periodic_stiffness(a, b, n, k, 0)
# sharded step: local deposit -> all-reduce(sum) of the grid -> replicated solve -> local gather/push
xs, vs_, ws = x[lo:hi].copy(), v[lo:hi].copy(), w[lo:hi]
dt = 0.1
for step in range(3):
    xs += 0.5 * dt * vs_
    part = torch.from_numpy(orc.deposit_periodic(xs, ws, a, b, n, k, 0))
    dist.all_reduce(part)                      # the ONLY exchange of the path
    phi = orc.poisson_solve(S, part.numpy())
    vs_ -= dt * orc.eval_dphi(xs, a, b, n, k, 0, phi)
    xs += 0.5 * dt * vs_
# unsharded reference on every rank
xo, vo = x.copy(), v.copy()
orc.integrate_vp(xo, vo, w, dt, 1.0, 3, 0, a, b, n, k, 0, S)
assert np.max(np.abs(xs - xo[lo:hi])) < 1e-12 and np.max(np.abs(vs_ - vo[lo:hi])) < 1e-12
# every rank must hold bit-identical phi (replicated solve)
t = torch.from_numpy(phi.copy()); ref = t.clone(); dist.broadcast(ref, src=0)
assert torch.equal(t, ref)
dist.barrier()
print("ok", rank)
"""


@pytest.mark.timeout(300)
def test_sharded_step_world2_gloo(tmp_path):
    """N>1 path on CPU: shard -> local deposit -> gloo all-reduce -> replicated solve == unsharded step."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


# ------------------------------------------------------------ launch plans (pure host logic, no GPU) ---
B200_SMS, B200_SMEM_OPTIN = 148, 232448


@pytest.mark.parametrize("order", [2, 3, 4, 5, 6])
def test_pass_plans_fit_the_device_for_every_mesh_size(vm, order):
    """vm_pass_plan_query over every mesh size the ABI accepts: the plan fits in shared memory, the CTA fits the
    launch bound of the kernel tier that runs it, the register budget of that bound covers the pipeline depth, and
    the replica grids + gather table are fully accounted for in the dynamic shared memory."""
    L = vm._lib
    sizes = list(range(1, 300)) + list(range(300, 4097, 37)) + [512, 1024, 2048, 4096]
    for n in sizes:
        for pass_ in (0, 1, 2):
            p = L.pass_plan(n, order, pass_)
            what = (n, order, pass_, p.variant, p.replicas, p.grid, p.threads, p.pairs, p.max_threads, p.gather_copies, p.smem_bytes)
            assert p.smem_bytes <= B200_SMEM_OPTIN, what
            ctas = p.grid // B200_SMS
            assert p.grid == ctas * B200_SMS and ctas in (1, 2), what          # whole waves of the SM count
            assert ctas * (p.smem_bytes + 1024) <= 227 * 1024 + 1024, what     # all CTAs of an SM resident together
            assert p.threads % 32 == 0 and 32 <= p.threads <= p.max_threads <= 1024, what
            assert ctas * p.threads <= 1024, what
            assert p.pairs in (1, 2, 4, 8), what
            regs = min(255, 65536 // max(p.max_threads, ctas * p.threads))       # what the launch bound lets ptxas use
            assert regs >= {1: 64, 2: 64, 4: 128, 8: 255}[p.pairs], what
            warps = p.threads // 32
            rows = n + order - 1
            grids = rows * p.replicas * 8 * (1 if p.variant in (2, 5) else warps)
            table = (n + order) * 8 * p.gather_copies if pass_ == 1 else 0
            assert p.smem_bytes >= grids + table + p.threads * 8, what
            if p.variant == 5:                # limb atomics: ONE two-limb grid per CTA, whose words double as the finish's work area
                assert p.replicas in (1, 2, 4, 8, 16, 32) and p.pairs == (2 if pass_ == 0 else 1) and p.max_threads == 1024, what
                pitch = rows + ((32 // p.replicas) % 32 - rows) % 32      # replica r is shifted by r * 32/R banks
                assert pitch >= rows and pitch % 32 == (32 // p.replicas) % 32, what
                assert p.smem_bytes >= max(pitch * p.replicas, 3 * n + 2) * 8 + table + p.threads * 8, what
                assert p.gather_copies in (1, 16) and (pass_ == 1 or p.gather_copies == 1), what
                assert n >= (88 if pass_ == 0 else 20), what
                continue
            if p.variant == 4:                # bank-sorted pass: one replica per warp + 32 class queues of 16 words per warp
                assert p.replicas == 1 and ctas == 1 and p.pairs == 1 and warps >= 4, what
                assert p.smem_bytes >= grids + table + p.threads * 8 + warps * 32 * 16 * 8, what
                assert p.gather_copies in (1, 2, 4, 8, 16) and (pass_ == 1 or p.gather_copies == 1), what
                assert p.max_threads == (512 if warps <= 16 else 1024), what
                continue
            if p.variant == 0:
                assert p.replicas == 32 and warps >= 6, what
            if p.variant == 3:
                assert p.replicas in (8, 16) and warps >= 24, what
            if p.gather_copies != 1:          # the 16-fold table exists only in the lane-private fused pass
                assert p.gather_copies == 16 and pass_ == 1 and p.variant == 0 and n > 16, what
            elif pass_ == 1 and p.variant == 0 and n > 16:   # ... and is only given up when it would cost the variant
                assert rows * 256 * 6 + (n + order) * 128 + 6 * 256 > 227 * 1024 - 1024, what
            if p.variant != 0:
                assert p.pairs == (2 if pass_ == 0 else 1) and p.max_threads == 1024, what


def test_pass_plans_of_the_benchmarked_meshes(vm):
    """The configurations measured in profiles/README.md section 6 (cubic splines, B200)."""
    L = vm._lib
    p = L.pass_plan(16, 4, 1)
    assert (p.variant, p.grid, p.threads, p.pairs, p.gather_copies) == (0, 296, 512, 1, 1)
    p = L.pass_plan(19, 4, 1)                             # the last mesh whose fused step runs lane-private (VM_AF_MIN_N = 20)
    assert (p.variant, p.grid, p.pairs, p.gather_copies) == (0, 296, 1, 16)
    p = L.pass_plan(64, 4, 2)                             # (the deep lane-private tiers of the mid-size meshes now only run with af = -1)
    assert p.variant == 5
    p = L.pass_plan(80, 4, 0)                             # deposit-only: lane-private while a plan exists (VM_AF_MIN_N_DEPOSIT = 88)
    assert (p.variant, p.grid, p.threads, p.pairs, p.max_threads) == (0, 148, 320, 4, 512)
    # limb atomics (variant 5) above: one full CTA per SM, 16-fold gather table, any mesh size (profiles/r02b_af_ab.txt)
    for n in (20, 32, 64, 128, 256, 512, 1024):
        p = L.pass_plan(n, 4, 1)
        assert (p.variant, p.grid, p.threads, p.pairs, p.max_threads) == (5, 148, 1024, 1, 1024), n
        # bank-steered replicas: conflict-free atomics (32 replicas) with the 16-fold gather table up to 512 cells; replicas go first
        assert (p.replicas, p.gather_copies) == ((32, 16) if n <= 256 else ((32, 1) if n <= 512 else (16, 1))), n
        assert p.smem_bytes <= 156 * 1024, n                      # (the rest of the 256 KB stays L1 for the streams)
    assert L.pass_plan(64, 4, 0).variant == 0 and L.pass_plan(128, 4, 0).variant == 5 and L.pass_plan(1024, 4, 0).variant == 5
    assert L.pass_plan(4096, 4, 1).variant == 5 and L.pass_plan(4096, 4, 1).gather_copies == 1     # 16 copies no longer fit
    assert L.pass_plan(16, 4, 0, 1).variant == 2          # VM_DEPOSIT_ATOMIC: the warp-aggregated A/B variant
    for bad in ((0, 4, 1), (16, 7, 1), (16, 4, 3), (5000, 4, 1)):
        with pytest.raises(vm.VMError):
            L.pass_plan(*bad)


def test_status_codes_mirror_the_header(vm):
    header = (ROOT / "include" / "vlasov_b200.h").read_text()
    body = header[header.index("typedef enum vm_status"):header.index("} vm_status;")]
    for name, val in re.findall(r"(VM_[A-Z_]+)\s*=\s*(\d+)", body):
        assert getattr(vm._lib, name) == int(val), name
    assert len(re.findall(r"VM_[A-Z_]+\s*=\s*\d+", body)) == 7


def test_plan_query_and_errors_need_no_device(vm):
    """Calls that must work on a machine without a GPU: the plan query, and error reporting with a NULL context."""
    with pytest.raises(vm.VMError) as ei:
        vm._lib.pass_plan(16, 9, 1)
    assert ei.value.code == vm._lib.VM_ERR_INVALID and "order" in str(ei.value)
    lib = vm.lib()
    assert lib.vm_ctx_destroy(None) == 0 and lib.vm_particles_destroy(None) == 0
    assert lib.vm_field_destroy(None) == 0 and lib.vm_vspline_destroy(None) == 0
    assert lib.vm_launch_count(None) == 0 and lib.vm_particles_size(None) == -1


def test_mirror_host_logic_without_a_device(vm):
    """Pure-Python pieces of the reference-API mirror: time-step counting, snapshot chunking and the example
    structs' defaults (src/examples/*.jl) with their mapping onto vm_particles_fill."""
    api = vm.api
    assert api._ntime((0.0, 20.0), 0.1) == 200 and api._ntime((0.0, 500.0), 1e-2) == 50000      # scripts' values
    assert api._chunks(200, 0) == [200] and api._chunks(0, 0) == [] and api._chunks(10, 3) == [3, 3, 3, 1]
    assert sum(api._chunks(12345, 100)) == 12345
    L = vm._lib
    assert api._fill_args(api.NormalDistribution()) == (L.VM_FILL_NORMAL, [0.0, 1.0])                       # normal.jl:4
    assert api._fill_args(api.BumpOnTail()) == (L.VM_FILL_BUMP_ON_TAIL, [0.03, 0.3, 0.1, 0.5, 4.5])          # bumpontail.jl:9
    assert api._fill_args(api.DoubleMaxwellian()) == (L.VM_FILL_DOUBLE_MAXWELLIAN, [-5.0, 5.0, 3.0])         # doublemaxwellian.jl:5
    assert api._fill_args(api.UniformDistribution()) == (L.VM_FILL_UNIFORM, [0.0, 1.0, -2.0, 2.0])           # uniform.jl:5
    assert api._fill_args(api.ShiftedNormalV()) == (L.VM_FILL_SHIFTED_NORMAL_V, [-5.0, 5.0, 2.0])            # shiftednormalv.jl:5
    assert api._fill_args(api.ShiftedUniformDistribution()) == (L.VM_FILL_SHIFTED_UNIFORM, [0.0, 1.0, -2.0, 2.0, 2.0])
    with pytest.raises(TypeError):
        api._fill_args(object())
    with pytest.raises(NotImplementedError):
        api.DiffEqIntegrator()
    b = api.PeriodicBasisBSplineKit((0.0, 1.0), 3, 16)
    assert (b.order, b.domain) == (3, (0.0, 1.0))


def test_header_is_plain_c_and_links_from_c(vm, tmp_path):
    """The boundary is a C ABI: include/vlasov_b200.h must compile as strict C11 (no C++-isms, no warnings), the
    example host program must link against libvlasov_b200.so without any Python or torch in the process, and
    without a GPU it must fail loudly with the library's own message (exit status 3)."""
    import shutil
    import torch
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pkg = ROOT / "vlasovmethods.jl_b200"
    exe = tmp_path / "landau"
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}",
           str(ROOT / "examples" / "landau_damping.c"), "-o", str(exe), f"-L{pkg}", "-lvlasov_b200", "-lm",
           f"-Wl,-rpath,{pkg}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        r = subprocess.run([str(exe), "200000", "32", "60"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "damping rate" in r.stdout, r.stdout[-500:] + r.stderr[-500:]
    else:
        r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
        assert r.returncode == 3 and "no CPU path" in r.stderr, (r.returncode, r.stderr)


def test_sobol_direction_numbers_of_the_device_load():
    """The 2-D Sobol construction vm_fill.cu uses for the bump-on-tail draws (fill kinds 7 / 8), restated in Python: point k =
    XOR over the set bits b of gray(k) of V_b, dimension 1 V_b = 2^(31-b), dimension 2 m_0 = 1, m_b = 2 m_(b-1) ^ m_(b-1),
    V_b = m_b 2^(31-b) (Joe-Kuo, primitive polynomial x + 1).  Must be scipy's unscrambled generator point for point (the
    GPU test compares the device output with the same generator)."""
    from scipy.stats import qmc

    def sobol2(k):
        g = k ^ (k >> 1)
        x2, m, b = 0, 1, 0
        while g >> b:
            if (g >> b) & 1:
                x2 ^= (m << (31 - b)) & 0xFFFFFFFF
            m = (m ^ (m << 1)) & 0xFFFFFFFF
            b += 1
        x1 = int(format(g, "032b")[::-1], 2)
        return x1 / 2.0**32, x2 / 2.0**32

    ref = qmc.Sobol(2, scramble=False, bits=32).random(1 << 12)
    mine = np.array([sobol2(k) for k in range(1 << 12)])
    assert np.array_equal(mine, ref)
    sob = qmc.Sobol(2, scramble=False, bits=32)
    sob.fast_forward((1 << 20) + 1)
    assert np.array_equal(np.array([sobol2((1 << 20) + 1 + j) for j in range(64)]), sob.random(64))
