"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI via
the Scattering1D frontend, against the numpy oracle, the committed golden outputs of the
live reference and the reference's own known-answer fixture."""
import os

import numpy as np
import pytest
import torch

from helpers import CONFIGS, GOLDEN, OVERSAMPLING, path_tolerance, rel_l2
from oracle.scattering1d_oracle import ScatteringOracle

pytestmark = pytest.mark.gpu

_mods = {}


def module_of(name):
    from tebscat import Scattering1D
    if name not in _mods:
        J, N, Q, T, mo = CONFIGS[name]
        _mods[name] = Scattering1D(J, N, Q, max_order=mo, T=T, oversampling=OVERSAMPLING.get(name, 0)).cuda()
    return _mods[name]


def gpu_forward(name, x):
    S, P = module_of(name)(torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda())
    torch.cuda.synchronize()
    return S.cpu().numpy(), P


@pytest.mark.parametrize('name', ['H', 'P', 'S', 'T', 'K'])
def test_parity_with_oracle(name):
    from tebscat.synth import ctg_batch, randn_batch
    J, N, Q, T, mo = CONFIGS[name]
    x = np.concatenate([randn_batch(4, N, 2, seed=4321).reshape(8, N).numpy(),
                        ctg_batch(4, N, seed=1234).reshape(8, N).numpy()])
    out, P = gpu_forward(name, x)
    ref64 = ScatteringOracle(J, N, Q, T, mo)(x)
    ref32 = ScatteringOracle(J, N, Q, T, mo, cdtype=np.complex64)(x)
    assert out.shape == ref64.shape and tuple(P.shape) == (16, 1) + ref64.shape[1:]
    err = np.linalg.norm(out.astype(np.float64) - ref64, axis=-1)
    tol = path_tolerance(ref64, ref32)
    assert np.all(err <= tol), float((err / tol).max())
    # randn rows: the north-star bound, rel-L2 <= 1e-5 for every coefficient path
    assert rel_l2(out[:8].astype(np.float64), ref64[:8], axis=-1).max() < 1e-5
    assert rel_l2(out.astype(np.float64), ref64) < 1e-6


def test_reference_known_answer_fixture():
    """kymatio/tests/scattering1d/test_torch_scattering1d.py:82-113 on test_data_1d.npz."""
    d = np.load(os.path.join(GOLDEN, 'kat_test_data_1d.npz'))
    out, _ = gpu_forward('K', d['x'])
    assert out.shape == d['Sx'].shape
    assert np.allclose(out, d['Sx'], rtol=1e-5, atol=1e-7)
    assert rel_l2(out, d['Sx'], axis=-1).max() < 1e-5


@pytest.mark.parametrize('name', ['H', 'P', 'S', 'T', 'O'])
def test_golden_reference_outputs(name):
    d = np.load(os.path.join(GOLDEN, 'scat_%s.npz' % name))
    out, _ = gpu_forward(name, d['x'])
    J, N, Q, T, mo = CONFIGS[name]
    ref64 = ScatteringOracle(J, N, Q, T, mo, oversampling=OVERSAMPLING.get(name, 0))(d['x'])
    tol = np.maximum(1e-5 * np.linalg.norm(ref64, axis=-1),
                     4.0 * np.linalg.norm(d['S'].astype(np.float64) - ref64, axis=-1))
    assert np.all(np.linalg.norm(out.astype(np.float64) - ref64, axis=-1) <= tol)
    assert rel_l2(out, d['S']) < 2e-6
    half = d['x'].shape[0] // 2                     # second half of the fixture batch is randn
    assert rel_l2(out[half:], d['S'][half:], axis=-1).max() < 1e-5


def test_simple_signals():
    """The disabled-but-valid properties of test_torch_scattering1d.py:47-77."""
    J, N, Q, T, mo = CONFIGS['H']
    S = module_of('H')
    meta = S.meta()
    z, _ = S(torch.zeros(2, N, device='cuda'))
    assert z.abs().max().item() == 0.0
    c, _ = S(torch.full((1, N), 0.7345, device='cuda'))
    assert c[:, 1:].abs().max().item() < 1e-6
    assert abs(c[0, 0].mean().item() - 0.7345) < 1e-5
    t = torch.arange(N, dtype=torch.float32)
    for k in (37, 401, 1500):
        s, _ = S(torch.cos(2 * np.pi * k * t / N)[None].cuda())
        assert s[:, torch.from_numpy(meta['order']) != 1, :].abs().max().item() < 1e-2


def test_batch_shape_agnostic_and_views():
    """test_torch_scattering1d.py:338-386: any leading batch dims, (N,) allowed."""
    J, N, Q, T, mo = CONFIGS['S']
    S = module_of('S')
    x = torch.randn(2, 3, N, device='cuda')
    a, P = S(x)
    assert a.shape == (2, 3, S.output_size(), S._sched[1].n_out) and P.shape[0] == 6
    b, _ = S(x.reshape(6, N))
    assert torch.equal(a.reshape(6, *a.shape[2:]), b)
    v, _ = S(x[0, 0])
    assert v.shape == a.shape[2:] and torch.equal(v, a[0, 0])
    e, _ = S(torch.zeros(0, N, device='cuda'))
    assert e.shape[0] == 0


@pytest.mark.parametrize('B', [2500, 6000])
def test_host_entry_point_matches_device_entry_point(B):
    """B = 2500: a few uniform chunks; B = 6000: the ramped chunk sizes of long batches (148, 296, 592 signals up and
    down around full chunks of 1184 and one partial chunk)."""
    J, N, Q, T, mo = CONFIGS['H']
    S = module_of('H')
    x = torch.randn(B, N, generator=torch.Generator().manual_seed(B)).pin_memory()
    h = S.scattering_host(x)
    d, _ = S(x.cuda())
    assert torch.equal(h, d.cpu())


def test_full_size_batch_properties():
    """BASELINE config 2: 8192 two-channel signals.  Size-independent checks: rows of a
    repeated signal are bit-identical wherever they sit in the batch, positive homogeneity
    S(a x) = a S(x) for a = 2^k (exact in binary floating point), and a sample of rows
    against the oracle."""
    from tebscat.synth import ctg_batch
    J, N, Q, T, mo = CONFIGS['H']
    S = module_of('H')
    base = ctg_batch(64, N, seed=99).reshape(128, N)
    x = base.repeat(128, 1).cuda()                 # 16384 signals
    out, _ = S(x)
    torch.cuda.synchronize()
    assert out.shape == (16384, 126, 75)
    assert torch.equal(out[:128], out[-128:]) and torch.equal(out[:128], out[8192:8320])
    scaled, _ = S(x[:256] * 4.0)
    assert torch.equal(scaled, out[:256] * 4.0)
    pick = [0, 77, 127]
    ref64 = ScatteringOracle(J, N, Q, T, mo)(base[pick].numpy())
    ref32 = ScatteringOracle(J, N, Q, T, mo, cdtype=np.complex64)(base[pick].numpy())
    got = out[pick].cpu().numpy().astype(np.float64)
    assert np.all(np.linalg.norm(got - ref64, axis=-1) <= path_tolerance(ref64, ref32))


def test_plan_validation_errors():
    """The C ABI refuses malformed plans with an error code and a message (never throws)."""
    import ctypes
    from tebscat import _lib
    from tebscat.schedule import build_plan
    from tebscat.torch_frontend import _DevicePlan
    J, N, Q, T, mo = CONFIGS['T']
    p = build_plan(J, N, Q, T, mo)
    p.tasks = p.tasks.copy()
    p.tasks[3, 1] = 5000                            # thread range outside the CTA
    with pytest.raises(ValueError) as ve:
        _DevicePlan(p, 0)
    assert 'outside the CTA' in str(ve.value)
    rc = _lib.load().tebscat_scat1d_forward(None, None, 1, None, None)
    assert rc == _lib.TEBSCAT_EINVAL


def test_plan_error_paths_free_the_plan_once():
    """Two error paths of tebscat_plan_create used to free the plan twice (round-1 advice): two tasks of one
    step on the same warps -> ValueError (TEBSCAT_EINVAL); dynamic + static shared memory above the opt-in
    limit -> NotImplementedError (TEBSCAT_EUNSUPPORTED).  The process survives and a valid plan still works."""
    from tebscat.schedule import build_plan
    from tebscat.torch_frontend import _DevicePlan
    J, N, Q, T, mo = CONFIGS['T']
    for _ in range(3):
        p = build_plan(J, N, Q, T, mo)
        p.tasks = p.tasks.copy()
        st = next(s for s in range(p.steps.shape[0]) if p.steps[s, 1] - p.steps[s, 0] >= 2)
        a, b = p.steps[st, 0], p.steps[st, 0] + 1
        p.tasks[b, 1] = p.tasks[a, 1]                  # second task starts on the first one's warps
        with pytest.raises(ValueError) as ve:
            _DevicePlan(p, 0)
        assert 'overlapping thread ranges' in str(ve.value)
        q = build_plan(J, N, Q, T, mo)
        optin = torch.cuda.get_device_properties(0).shared_memory_per_block_optin
        q.smem_complex = optin // 8 - 204              # dynamic part alone fits exactly; with the static part it does not
        with pytest.raises(NotImplementedError) as ne:
            _DevicePlan(q, 0)
        assert 'shared memory' in str(ne.value)
        q.smem_complex = optin // 8 + 4096             # and the plainly oversized request
        with pytest.raises(NotImplementedError):
            _DevicePlan(q, 0)
    from tebscat import Scattering1D
    S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    x = np.random.RandomState(5).randn(2, N).astype(np.float32)
    out, _ = S(torch.from_numpy(x).cuda())
    ref = ScatteringOracle(J, N, Q, T, mo)(x)
    assert rel_l2(out.cpu().numpy().astype(np.float64), ref) < 1e-5


@pytest.mark.parametrize('cfg', [(7, 8, 512, 128, 2), (8, 8, 1000, 256, 2), (10, 4, 5000, 1024, 1), (5, 8, 512, 24, 2),
                                 (3, 8, 2048, 6, 2), (10, 12, 5000, 768, 2)])
def test_other_configurations(cfg):
    """Wide folds (k > 128), output-rate lengths of 2..8 samples (tiny transforms), T < 2**J,
    non power-of-two T, unbatchable buffers."""
    from tebscat import Scattering1D
    J, Q, N, T, mo = cfg
    S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    x = np.random.RandomState(3).randn(3, N).astype(np.float32)
    out, _ = S(torch.from_numpy(x).cuda())
    out = out.cpu().numpy().astype(np.float64)
    ref = ScatteringOracle(J, N, Q, T, mo)(x)
    assert out.shape == ref.shape
    nr = np.linalg.norm(ref, axis=-1)
    err = np.linalg.norm(out - ref, axis=-1)
    assert np.all(err <= 1e-5 * nr + 1e-10 * nr.max()), float((err / nr).max())


def test_output_conventions_list_and_dict():
    """out_type='list' and vectorize=False (core/scattering1d.py:379-384,
    frontend/torch_frontend.py:240-253; test_torch_scattering1d.py:195-243 test_coordinates):
    same numbers as the vectorised array, keyed / ordered by meta()['key']."""
    from tebscat import Scattering1D
    S = Scattering1D(4, 700, 2, T=16).cuda()
    x = torch.randn(3, 2, 700, device='cuda')
    arr, _ = S(x)
    meta = S.meta()
    S.out_type = 'list'
    lst, _ = S(x)
    assert isinstance(lst, list) and len(lst) == arr.shape[-2] and set(lst[0]) == {'coef', 'j'}
    for c, item in enumerate(lst):
        assert item['coef'].shape == (3, 2, arr.shape[-1]) and torch.equal(item['coef'], arr[..., c, :])
        assert len(item['j']) == meta['order'][c]
    S.out_type = 'array'
    S.vectorize = False
    with pytest.warns(DeprecationWarning):
        dct, _ = S(x)
    assert list(dct.keys()) == meta['key']
    for c, k in enumerate(meta['key']):
        assert dct[k].shape == (3, 2, 1, arr.shape[-1]) and torch.equal(dct[k][..., 0, :], arr[..., c, :])


@pytest.mark.gpu
def test_relaxed_and_chained_schedule_is_bitwise_equal_to_fully_barriered(monkeypatch):
    """Race check without a sanitizer: warp-fenced chained passes and relaxed barriers change WHEN a butterfly
    runs, never its arithmetic, so the product schedule must reproduce -- bit for bit, run after run -- the
    output of the same cascade scheduled with one CTA barrier after every single pass."""
    import torch
    from tebscat import Scattering1D
    from tebscat.synth import ctg_batch
    J, N, Q, T = 6, 4800, 8, 64
    x = ctg_batch(148 * 2, N, seed=11).reshape(-1, N)[:148 * 3].cuda()
    fast = Scattering1D(J, N, Q, T=T).cuda()
    ref = fast(x)[0].clone()
    for _ in range(4):                                      # run-to-run determinism
        assert torch.equal(fast(x)[0], ref)
    monkeypatch.setenv('TEBSCAT_RELAX', '0')
    monkeypatch.setenv('TEBSCAT_CHAIN', '0')
    safe = Scattering1D(J, N, Q, T=T).cuda()
    assert safe._schedule().stats['n_relaxed'] == 0 and safe._schedule().stats['n_steps'] > fast._schedule().stats['n_steps']
    assert torch.equal(safe(x)[0], ref)


def test_forward_is_reentrant_across_host_threads_and_streams():
    """SURVEY 8b 'Threading': a plan is immutable after creation, so forward calls from several host threads on
    different streams must not interfere."""
    import threading
    from tebscat import Scattering1D
    S = Scattering1D(6, 4800, 8, T=64).cuda()
    xs = [torch.randn(300 + 50 * i, 4800, generator=torch.Generator().manual_seed(i)).cuda() for i in range(4)]
    serial = [S(x)[0].clone() for x in xs]
    torch.cuda.synchronize()
    out = [None] * 4

    def work(i):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(3):
                out[i] = S(xs[i])[0]
        st.synchronize()
    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for a, b in zip(serial, out):
        assert torch.equal(a, b)


@pytest.mark.parametrize('layout', ['shared', 'scratch'])
@pytest.mark.parametrize('cfg', [(4, 4827, 8, 4, 2), (3, 4223, 1, 4, 1), (2, 4085, 1, 2, 2), (4, 4827, 8, 1, 2)])
def test_configurations_beyond_the_fused_schedule(cfg, layout, monkeypatch):
    """Output-rate lengths of 2048 samples and more (T <= 4 at a padded length of 8192) do not fit the fused
    single-kernel schedule while U0 lives in shared memory (build_plan says so); the frontend then serves them on the
    op-by-op CUDA level of tebscat/large.py -- same parity bar, forward and backward.  With U0 parked in the global
    scratch (the default layout) lengths of 2048 and 4096 fit the fused kernel; 8192 (T = 1) stays op by op."""
    from tebscat import Scattering1D
    from tebscat.schedule import build_plan
    from oracle.scattering1d_grad_oracle import GradOracle
    J, N, Q, T, mo = cfg
    monkeypatch.setenv('TEBSCAT_U0_GLOBAL', '0' if layout == 'shared' else '1')
    fused = layout == 'scratch' and T > 1
    if not fused:
        with pytest.raises(NotImplementedError):
            build_plan(J, N, Q, T, mo)
    S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    x = torch.randn(3, N, generator=torch.Generator().manual_seed(J)).cuda().requires_grad_(True)
    out, _ = S(x)
    assert S._op_by_op == (not fused)
    ref = ScatteringOracle(J, N, Q, T, mo)(x.detach().cpu().numpy())
    got = out.detach().cpu().numpy().astype(np.float64)
    assert got.shape == ref.shape
    nr = np.linalg.norm(ref, axis=-1)
    assert np.all(np.linalg.norm(got - ref, axis=-1) <= 1e-5 * nr + 1e-10 * nr.max())
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(9))
    (out * w.cuda()).sum().backward()
    _, g64 = GradOracle(J, N, Q, T, mo).vjp(x.detach().cpu().numpy(), w.numpy())
    assert rel_l2(x.grad.cpu().numpy(), g64, axis=-1).max() < 1e-5


@pytest.mark.parametrize('name', ['H', 'P', 'K'])
def test_u0_layouts_agree_bit_for_bit(name, monkeypatch):
    """Where the consumers of U0 find it -- a per-CTA global scratch (default: OP_STOREC after the root transform,
    OP_GMULFOLD / OP_GMULFOLD2 multiplies) or shared memory (round 1) -- changes the schedule, never the arithmetic:
    same bits, on a batch of several signals per CTA (the scratch is rewritten for every signal), eagerly, on two
    streams at once (one scratch per stream) and replayed from a CUDA graph (capture-time scratch)."""
    from tebscat import Scattering1D
    from helpers import CONFIGS
    J, N, Q, T, mo = CONFIGS[name]
    x = torch.randn(148 * 3 + 5, N, generator=torch.Generator().manual_seed(3)).cuda()
    S1 = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    a = S1(x)[0].clone()
    assert S1._schedule().scratch_complex == 1 << S1.J_pad
    monkeypatch.setenv('TEBSCAT_U0_GLOBAL', '0')
    S0 = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    assert S0._schedule().scratch_complex == 0
    assert torch.equal(S0(x)[0], a)
    # two streams at once
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        b1 = [S1(x)[0] for _ in range(3)]
    with torch.cuda.stream(s2):
        b2 = [S1(x.flip(0))[0] for _ in range(3)]
    torch.cuda.synchronize()
    assert all(torch.equal(b, a) for b in b1) and all(torch.equal(b, a.flip(0)) for b in b2)
    # CUDA graph
    xs = x.clone()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        S1(xs)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        og = S1(xs)[0]
    xs.copy_(x.flip(0))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(og, a.flip(0))


@pytest.mark.gpu
@pytest.mark.parametrize('cfg', [('T', None), ('T', 1), ('L', None)])
def test_analysis_window_on_every_level(cfg):
    """Scattering1D.set_window: the taper the kernels' loads apply (the Tukey option of the phase module,
    kymatio_phase_scattering.py:405-407) equals tapering the input first -- on the fused level, the op-by-op level
    (forced through a configuration the fused schedule cannot hold) and the large-support level, forward and
    backward (gradient of a windowed transform = window * gradient)."""
    from tebscat import Scattering1D
    name, force_T = cfg
    if name == 'L':
        J, N, Q, T, mo = 6, 9000, 4, 64, 2
    else:
        J, N, Q, T, mo = CONFIGS[name]
    if force_T is not None:
        J, N, Q, T, mo = 6, 4800, 8, 1, 1                 # output rate 8192 at Np = 8192 (no averaging): op-by-op level
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, N, generator=g).cuda()
    w = torch.linspace(0.2, 1.0, N).cuda() * torch.cos(torch.linspace(0, 3.0, N)).cuda().abs()
    plain = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    tapered = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    tapered.set_window(w.cpu().numpy())
    assert tapered._op_by_op == (force_T is not None)
    ref, _ = plain((x * w).contiguous())
    out, _ = tapered(x)
    assert torch.equal(out, ref)
    out2, _ = tapered(x)                                   # graph replay on the large levels
    assert torch.equal(out2, ref)
    xg = x.clone().requires_grad_(True)
    wgt = torch.randn(ref.shape, generator=torch.Generator().manual_seed(12)).cuda()
    (tapered(xg)[0] * wgt).sum().backward()
    xr = (x * w).contiguous().requires_grad_(True)
    (plain(xr)[0] * wgt).sum().backward()
    assert torch.allclose(xg.grad, xr.grad * w, rtol=1e-6, atol=1e-6 * float(xr.grad.abs().max()))
    tapered.set_window(None)
    assert torch.equal(tapered(x)[0], plain(x)[0])


@pytest.mark.gpu
def test_c_consumer_without_python_at_run_time(tmp_path):
    """The ABI is self-sufficient for a non-Python consumer: examples/c_consumer.c, plain C compiled with
    gcc against include/tebscat.h, loads a plan FILE and transforms host signals; same numbers as the torch frontend,
    bit for bit."""
    import subprocess
    from tebscat import Scattering1D
    from tebscat.export_plan import main as export_main
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, 'vae-teb_b200', 'tebscat')
    exe = str(tmp_path / 'c_consumer')
    subprocess.run(['gcc', '-O2', '-I', os.path.join(root, 'include'), os.path.join(root, 'examples', 'c_consumer.c'), '-o', exe,
                    '-L', libdir, '-ltebscat', '-Wl,-rpath,' + libdir, '-lm'], check=True)
    J, N, Q, T, mo = CONFIGS['S']
    plan_path, x_path, s_path = (str(tmp_path / n) for n in ('S.tebplan', 'x.f32', 'S.f32'))
    export_main(['--J', str(J), '--shape', str(N), '--Q', str(Q), '--T', str(T), '--max-order', str(mo), plan_path])
    x = np.random.RandomState(8).randn(37, N).astype(np.float32)
    x.tofile(x_path)
    r = subprocess.run([exe, plan_path, x_path, s_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    S = Scattering1D(J, N, Q, max_order=mo, T=T).cuda()
    ref = S(torch.from_numpy(x).cuda())[0].cpu().numpy()
    got = np.fromfile(s_path, np.float32).reshape(ref.shape)
    assert np.array_equal(got, ref)
    bad = subprocess.run([exe, x_path, x_path, s_path], capture_output=True, text=True)      # not a plan file
    assert bad.returncode == 1 and 'not a tebscat plan file' in bad.stderr
