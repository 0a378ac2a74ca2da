"""GPU tier: the exchange + sharded-Adam kernel of ray-sharded training (csrc/optim.cu `k_adam_step_p2p`) on ONE
GPU.  The kernel only sees per-rank device POINTERS (gradient, fp16 shadow, flag buffer of every rank), so R local buffers in
the pointer arrays emulate R ranks exactly: rank r's launch reads slice r of all R gradients, updates its fp32 master slice
and writes the new fp16 values into all R shadow tables.  Checked against `lnrf_adam_step` (itself checked against
torch.optim.Adam + GradScaler in test_gpu_fused.py) on the fp32 mean of the R fp16 gradients:

  * result == single-GPU Adam on the averaged gradient, over several steps, R = 2 and R = 8;
  * every "rank" ends with the same fp16 table, equal to half(master);
  * a non-finite flag raised by ANY rank skips the update everywhere and reports found_inf;
  * the in-kernel synchronisation variant (`lnrf_adam_step_sharded_sync` + `lnrf_exchange_finish`): R launches on R streams
    meet inside the kernels, same numbers, gradients cleared, epochs advance.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HYPER = (1e-2, 0.9, 0.99, 1e-15, 0.0)


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


class _Ranks:
    """R emulated ranks: full fp16 gradient / shadow / flag buffer each, fp32 master + moments for the owned slice only."""

    def __init__(self, dev, R, Sz, seed=0):
        from laenerf_b200 import _native as N
        self.N, self.R, self.Sz, self.P = N, R, Sz, R * Sz
        g = torch.Generator(device=dev).manual_seed(seed)
        self.p0 = (torch.rand(self.P, device=dev, generator=g) - 0.5)
        self.grad = [torch.zeros(self.P, dtype=torch.half, device=dev) for _ in range(R)]
        self.shadow = [self.p0.half() for _ in range(R)]
        self.flag = [torch.zeros(32, dtype=torch.float32, device=dev) for _ in range(R)]
        self.master = [self.p0[r * Sz:(r + 1) * Sz].clone() for r in range(R)]
        self.m = [torch.zeros(Sz, device=dev) for _ in range(R)]
        self.v = [torch.zeros(Sz, device=dev) for _ in range(R)]
        self.scale = torch.full((1,), 1024.0, device=dev)
        self.step_count = torch.ones(1, device=dev)
        self.found = [torch.zeros(1, device=dev) for _ in range(R)]
        self.sync_state = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in range(R)]
        mk = lambda ts: (C.c_void_p * R)(*[t.data_ptr() for t in ts])
        self.arrays = (mk(self.grad), mk(self.shadow), mk(self.flag))

    def fill_grads(self, seed, bad_rank=None):
        dev = self.p0.device
        g = torch.Generator(device=dev).manual_seed(seed)
        for r in range(self.R):
            self.grad[r].copy_((torch.randn(self.P, device=dev, generator=g) * 8.0).half())  # "scaled" gradients, |g| ~ 8
            self.flag[r][0] = 0.0
        if bad_rank is not None:
            self.grad[bad_rank][12345 % self.P] = float("inf")
        # every rank checks its own gradient and publishes the flag beside it (optim.py _step_sharded)
        N = self.N
        for r in range(self.R):
            arr = (N.OptTensor * 1)()
            arr[0].grad, arr[0].n, arr[0].grad_dtype = self.grad[r].data_ptr(), self.P, N.F16
            N.check(N.lib().lnrf_grad_nonfinite_check(C.cast(arr, C.c_void_p), 1, N.ptr(self.flag[r]), None))

    def mean_grad_f32(self):
        acc = torch.zeros(self.P, device=self.p0.device)
        for r in range(self.R):  # the kernel's order: rank 0, 1, ... added in fp32, then one multiply by 1/R
            acc += self.grad[r].float()
        return acc * (1.0 / self.R)

    def launch(self, r, stream=None, sync=False):
        N = self.N
        g, sh, fl = self.arrays
        st = None if stream is None else stream.cuda_stream
        if sync:
            N.check(N.lib().lnrf_adam_step_sharded_sync(C.cast(g, C.c_void_p), C.cast(sh, C.c_void_p), C.cast(fl, C.c_void_p), self.R, r,
                                                        r * self.Sz, self.Sz, N.ptr(self.master[r]), N.ptr(self.m[r]), N.ptr(self.v[r]), *HYPER,
                                                        N.ptr(self.scale), N.ptr(self.found[r]), N.ptr(self.step_count), None,
                                                        N.ptr(self.sync_state[r]), st))
        else:
            N.check(N.lib().lnrf_adam_step_sharded(C.cast(g, C.c_void_p), C.cast(sh, C.c_void_p), C.cast(fl, C.c_void_p), self.R, r * self.Sz,
                                                   self.Sz, N.ptr(self.master[r]), N.ptr(self.m[r]), N.ptr(self.v[r]), *HYPER, N.ptr(self.scale),
                                                   N.ptr(self.found[r]), N.ptr(self.step_count), None, st))


class _Single:
    """lnrf_adam_step on one full fp32 vector with an fp32 gradient: the comparison target."""

    def __init__(self, ranks):
        self.N = ranks.N
        self.p = ranks.p0.clone()
        self.m, self.v = torch.zeros_like(self.p), torch.zeros_like(self.p)
        self.p16 = self.p.half()
        self.scale, self.step_count = ranks.scale, ranks.step_count
        self.found = torch.zeros(1, device=self.p.device)

    def step(self, grad_f32):
        N = self.N
        g = grad_f32.clone()
        arr = (N.OptTensor * 1)()
        arr[0].params, arr[0].exp_avg, arr[0].exp_avg_sq = self.p.data_ptr(), self.m.data_ptr(), self.v.data_ptr()
        arr[0].grad, arr[0].params_f16, arr[0].n, arr[0].grad_dtype = g.data_ptr(), self.p16.data_ptr(), self.p.numel(), N.F32
        N.check(N.lib().lnrf_adam_step(C.cast(arr, C.c_void_p), 1, *HYPER, N.ptr(self.scale), N.ptr(self.found), N.ptr(self.step_count), None, None))
        torch.cuda.synchronize()


@pytest.mark.parametrize("R", [2, 8])
def test_sharded_exchange_adam_equals_single_gpu_adam_on_the_mean_gradient(dev, R):
    Sz = 8 * 1000 * 13  # a multiple of 8 that is not a multiple of the 2048-element chunk
    ranks = _Ranks(dev, R, Sz, seed=R)
    single = _Single(ranks)
    for it in range(4):
        ranks.fill_grads(100 + it)
        mean = ranks.mean_grad_f32()
        for r in range(R):
            ranks.launch(r)
        torch.cuda.synchronize()
        single.step(mean)
        ranks.step_count += 1.0  # lnrf_amp_update's job in the product path
        master = torch.cat(ranks.master)
        assert torch.allclose(master, single.p, rtol=0, atol=2e-7), float((master - single.p).abs().max())
        assert torch.allclose(torch.cat(ranks.m), single.m, rtol=1e-6, atol=1e-9)
        assert torch.allclose(torch.cat(ranks.v), single.v, rtol=1e-6, atol=1e-12)
        for r in range(R):
            assert torch.equal(ranks.shadow[r], ranks.shadow[0])            # replicas identical
            assert float(ranks.found[r]) == 0.0
        assert torch.equal(ranks.shadow[0], master.half())                  # the table every rank gathers from next step
        assert float((master - ranks.p0).abs().max()) > 1e-3                # it did move


def test_nonfinite_gradient_on_any_rank_skips_the_update_everywhere(dev):
    R, Sz = 4, 8 * 4096
    ranks = _Ranks(dev, R, Sz, seed=5)
    ranks.fill_grads(7, bad_rank=2)
    assert [float(f[0]) for f in ranks.flag] == [0.0, 0.0, 1.0, 0.0]
    before = [t.clone() for t in ranks.master]
    shadow_before = ranks.shadow[1].clone()
    for r in range(R):
        ranks.launch(r)
    torch.cuda.synchronize()
    for r in range(R):
        assert float(ranks.found[r]) == 1.0           # every rank takes the same decision
        assert torch.equal(ranks.master[r], before[r]) and float(ranks.m[r].abs().sum()) == 0.0
        assert torch.equal(ranks.shadow[r], shadow_before)
    # the next step with clean gradients proceeds normally
    ranks.fill_grads(8)
    for r in range(R):
        ranks.launch(r)
    torch.cuda.synchronize()
    assert all(float(f) == 0.0 for f in ranks.found)
    assert float((torch.cat(ranks.master) - ranks.p0).abs().max()) > 1e-3


@pytest.mark.parametrize("R", [2, 8])
def test_in_kernel_synchronisation_variant_with_emulated_peers(dev, R):
    """lnrf_adam_step_sharded_sync + lnrf_exchange_finish: the R launches must be co-resident (each waits for the others' arrive
    words), so they go to R streams; slices are small enough for all blocks to be resident at once."""
    from laenerf_b200 import _native as N
    Sz = 8 * 2048  # 8 blocks per launch
    ranks = _Ranks(dev, R, Sz, seed=20 + R)
    single = _Single(ranks)
    streams = [torch.cuda.Stream() for _ in range(R)]
    for it in range(3):
        ranks.fill_grads(300 + it)
        mean = ranks.mean_grad_f32()
        torch.cuda.synchronize()
        for r in range(R):
            with torch.cuda.stream(streams[r]):
                ranks.launch(r, streams[r], sync=True)
        for r in range(R):
            with torch.cuda.stream(streams[r]):
                N.check(N.lib().lnrf_exchange_finish(N.ptr(ranks.flag[r]), R, N.ptr(ranks.sync_state[r]), N.ptr(ranks.grad[r]), ranks.P,
                                                     streams[r].cuda_stream))
        torch.cuda.synchronize()
        single.step(mean)
        ranks.step_count += 1.0
        master = torch.cat(ranks.master)
        assert torch.allclose(master, single.p, rtol=0, atol=2e-7)
        for r in range(R):
            assert torch.equal(ranks.shadow[r], master.half())
            assert int(ranks.sync_state[r][0]) == it + 1 and int(ranks.sync_state[r][1]) == 0   # epoch advanced, ticket re-armed
            assert float(ranks.grad[r].float().abs().sum()) == 0.0                               # cleared by the closing kernel
            words = ranks.flag[r].view(torch.int32)
            assert words[8:8 + R].tolist() == [it + 1] * R and words[16:16 + R].tolist() == [it + 1] * R


def test_pipelined_exchange_with_two_gradient_buffers_equals_the_one_buffer_sequence(dev):
    """lnrf_grad_nonfinite_check_snapshot + lnrf_adam_step_sharded_pipelined (software-pipelined sharded step: two gradient buffers and
    two flag words alternate, the kernel clears the OTHER buffer / flag and performs GradScaler.update() itself) against the one-buffer
    sequence (check, lnrf_adam_step_sharded, lnrf_exchange_tail): same masters, moments, tables, scale, step number and growth tracker
    over six steps with a non-finite gradient in step 2 (skip + back-off) and a growth interval of 2; the buffer of the previous step
    is zero again after every step."""
    from laenerf_b200 import _native as N
    R, Sz = 4, 8 * 3000
    lib = N.lib()
    a = _Ranks(dev, R, Sz, seed=11)   # one buffer, closing launch
    b = _Ranks(dev, R, Sz, seed=11)   # two buffers, pipelined
    P = a.P
    b.grad2 = [[b.grad[r], torch.zeros(P, dtype=torch.half, device=dev)] for r in range(R)]
    mk = lambda ts: (C.c_void_p * R)(*[t if isinstance(t, int) else t.data_ptr() for t in ts])
    b.arr2 = [(mk([b.grad2[r][ph] for r in range(R)]), mk(b.shadow), mk([b.flag[r].data_ptr() + 16 * ph for r in range(R)])) for ph in range(2)]
    # per-"rank" scaler state (every rank keeps its own copy; they stay identical)
    st = {}
    for name, ranks in (("a", a), ("b", b)):
        st[name] = dict(scale=[torch.full((1,), 1024.0, device=dev) for _ in range(R)], tracker=[torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(R)],
                        step=[torch.ones(1, device=dev) for _ in range(R)], found=[torch.zeros(1, device=dev) for _ in range(R)],
                        snap=[torch.zeros(8, device=dev) for _ in range(R)])
    amp = (2.0, 0.5, 2)   # growth, back-off, interval

    def one(grad):
        arr = (N.OptTensor * 1)()
        arr[0].grad, arr[0].n, arr[0].grad_dtype = grad.data_ptr(), P, N.F16
        return arr

    for it in range(6):
        ph = it & 1
        gen = torch.Generator(device=dev).manual_seed(500 + it)
        grads = [(torch.randn(P, device=dev, generator=gen) * 8.0).half() for _ in range(R)]
        if it == 2:
            grads[1][777] = float("nan")
        # ---- one-buffer sequence
        for r in range(R):
            a.grad[r].copy_(grads[r])
            a.flag[r].zero_()
            N.check(lib.lnrf_grad_nonfinite_check(C.cast(one(a.grad[r]), C.c_void_p), 1, N.ptr(a.flag[r]), None))
        g, sh, fl = a.arrays
        for r in range(R):
            N.check(lib.lnrf_adam_step_sharded(C.cast(g, C.c_void_p), C.cast(sh, C.c_void_p), C.cast(fl, C.c_void_p), R, r * Sz, Sz, N.ptr(a.master[r]),
                                               N.ptr(a.m[r]), N.ptr(a.v[r]), *HYPER, N.ptr(st["a"]["scale"][r]), N.ptr(st["a"]["found"][r]),
                                               N.ptr(st["a"]["step"][r]), None, None))
        for r in range(R):
            N.check(lib.lnrf_exchange_tail(N.ptr(a.grad[r]), P, N.ptr(st["a"]["scale"][r]), N.ptr(st["a"]["tracker"][r]), N.ptr(st["a"]["found"][r]),
                                           N.ptr(st["a"]["step"][r]), *amp, None))
        # ---- pipelined: accumulate into buffer `ph`; the kernel clears buffer 1 - ph and flag word 1 - ph
        for r in range(R):
            assert float(b.grad2[r][ph].float().abs().max()) == 0.0      # cleared by the previous step's kernel (or never used)
            b.grad2[r][ph].copy_(grads[r])
            N.check(lib.lnrf_grad_nonfinite_check_snapshot(C.cast(one(b.grad2[r][ph]), C.c_void_p), 1, b.flag[r].data_ptr() + 16 * ph,
                                                           N.ptr(st["b"]["scale"][r]), N.ptr(st["b"]["step"][r]), N.ptr(st["b"]["snap"][r]), None))
        g, sh, fl = b.arr2[ph]
        for r in range(R):
            N.check(lib.lnrf_adam_step_sharded_pipelined(
                C.cast(g, C.c_void_p), C.cast(sh, C.c_void_p), C.cast(fl, C.c_void_p), R, r * Sz, Sz, N.ptr(b.master[r]), N.ptr(b.m[r]), N.ptr(b.v[r]),
                *HYPER, N.ptr(st["b"]["snap"][r]), None, b.grad2[r][1 - ph].data_ptr(), P, b.flag[r].data_ptr() + 16 * (1 - ph),
                N.ptr(st["b"]["scale"][r]), N.ptr(st["b"]["tracker"][r]), N.ptr(st["b"]["found"][r]), N.ptr(st["b"]["step"][r]), *amp, None))
        torch.cuda.synchronize()
        for r in range(R):
            for x, y in ((a.master[r], b.master[r]), (a.m[r], b.m[r]), (a.v[r], b.v[r]), (a.shadow[r], b.shadow[r])):
                assert torch.equal(x, y), (it, r)
            for k in ("scale", "tracker", "step", "found"):
                assert torch.equal(st["a"][k][r], st["b"][k][r]), (it, r, k, st["a"][k][r], st["b"][k][r])
            assert float(b.grad2[r][1 - ph].float().abs().max()) == 0.0 and float(b.flag[r][4 * (1 - ph)]) == 0.0
    assert float(st["b"]["step"][0]) == 6.0      # five applied steps + 1 (the NaN step was skipped)
    assert float(st["b"]["scale"][0]) == 1024.0 * 2.0 * 0.5 * 2.0   # grew after steps 0-1, backed off at 2, grew again after 3-4
