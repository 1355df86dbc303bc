"""Multi-GPU: ray shards on several devices, partial images summed onto the root — with the peer-memory kernel
that fuses reduce-scatter, fp16 finalize and gather (default) and with ncclReduce (fallback).
Needs >= 2 visible GPUs (gpurun --gpus 2); skipped otherwise."""
import ctypes as C
import threading

import numpy as np
import pytest

from light_garden_b200 import abi
from util import have_cuda, small_specs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="no CUDA device")]

NCCL, PEER = 1, 2   # lg_reduce_mode_set


def device_count():
    from light_garden_b200 import _lib
    n = C.c_int32(0)
    _lib.load().lg_device_count(C.byref(n))
    return n.value


def reduce_all(ctxs, root=0):
    errs = []

    def run(i):
        try:
            ms = C.c_float()
            ctxs[i].call("lg_image_reduce", root, C.byref(ms))
        except Exception as e:   # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=run, args=(i,)) for i in range(len(ctxs))]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errs, errs


@pytest.mark.skipif(not have_cuda() or device_count() < 2, reason="needs 2 GPUs")
def test_two_device_render_equals_one_device(oracle):
    from light_garden_b200 import _lib
    from light_garden_b200.tracer import Context, Renderer, Tracer
    spec = small_specs()["C1"]
    world = min(device_count(), 4)
    ctxs = [Context(d, abi.LG_PRECISION_F32) for d in range(world)]
    one = Context(0, abi.LG_PRECISION_F32)
    try:
        arr = (C.c_void_p * world)(*[c.h for c in ctxs])
        _lib.check(ctxs[0].h, _lib.load().lg_comm_init_all(arr, world))
        t1 = spec.apply(Tracer(spec.canvas_bounds, ctx=one))
        r1 = Renderer(one, spec.width, spec.height)
        r1.render(t1)
        ref = r1.read_rgba32f()
        tracers, rends = [], []
        for rk, c in enumerate(ctxs):
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=c))
            t.set_shard(rk, world)
            tracers.append(t)
            rends.append(Renderer(c, spec.width, spec.height))
        results = {}
        for mode in (PEER, NCCL, PEER):
            for rk, c in enumerate(ctxs):
                c.call("lg_reduce_mode_set", mode)
                rends[rk].clear(1.0 if rk == 0 else 0.0)     # only the root owns the clear alpha
                rends[rk].render(tracers[rk])
            parts = [r.read_rgba32f() for r in rends] if mode == PEER else None
            reduce_all(ctxs, 0)
            total = rends[0].read_rgba32f()
            half = rends[0].read_rgba16f()
            # same fragments, different fp32 summation trees (partial images + reduce vs one image): hot pixels next
            # to a light sum ~1e4 fragments, so the stated bound is 1e-4 relative to the pixel value
            assert np.array_equal(total[..., 3] > 1, ref[..., 3] > 1)
            rel = np.abs(total - ref) / np.maximum(1.0, np.abs(ref))
            assert rel.max() < 1e-4, (mode, rel.max())
            assert np.median(rel) < 1e-7
            # the Rgba16Float frame is the rounded fp32 sum, whichever kernel produced it
            assert np.array_equal(half.view(np.uint16), total.astype(np.float16).view(np.uint16)), mode
            if mode == PEER:   # deterministic: partial images added in rank order
                acc = parts[0].copy()
                for p in parts[1:]:
                    acc += p
                assert np.array_equal(total, acc)
            results[mode] = total
        assert (np.abs(results[PEER] - results[NCCL]) <= 1e-5 * np.maximum(1.0, np.abs(ref))).all()
    finally:
        one.close()
        for c in ctxs:
            c.close()


@pytest.mark.skipif(not have_cuda() or device_count() < 2, reason="needs 2 GPUs")
def test_reduce_survives_a_resized_image_on_the_device_side_barriers(oracle):
    """The steady state of lg_image_reduce has no NCCL call: two device-side barriers over peer-mapped flag words around
    the fused kernel.  When a rank's image moves (lg_image_configure with another size) its "exchange the handles again"
    bit rides on the first barrier, every rank skips the kernel, the handles travel through NCCL once more and the reduce
    is repeated -- also when only ONE rank noticed."""
    from light_garden_b200 import _lib
    from light_garden_b200.tracer import Context, Renderer, Tracer
    spec = small_specs()["C1"]
    ctxs = [Context(d, abi.LG_PRECISION_F32) for d in range(2)]
    try:
        arr = (C.c_void_p * 2)(*[c.h for c in ctxs])
        _lib.check(ctxs[0].h, _lib.load().lg_comm_init_all(arr, 2))
        tracers = []
        for rk, c in enumerate(ctxs):
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=c))
            t.set_shard(rk, 2)
            tracers.append(t)

        def frame(sizes):
            rends = [Renderer(c, w, h) for c, (w, h) in zip(ctxs, sizes)]      # lg_image_configure: buffers may move
            for rk, r in enumerate(rends):
                r.clear(1.0 if rk == 0 else 0.0)
                r.render(tracers[rk])
            parts = [r.read_rgba32f() for r in rends]
            reduce_all(ctxs, 0)
            total = rends[0].read_rgba32f()
            assert np.array_equal(total, parts[0] + parts[1])
            half = rends[0].read_rgba16f()
            assert np.array_equal(half.view(np.uint16), total.astype(np.float16).view(np.uint16))
            return total

        a = frame([(480, 270)] * 2)
        b = frame([(480, 270)] * 2)                    # steady state: flags only
        assert np.array_equal(a, b) or (np.abs(a - b) <= 1e-5 * np.maximum(1.0, np.abs(a))).all()
        c_ = frame([(640, 360)] * 2)                   # every rank's image moved
        assert c_.shape == (360, 640, 4)
        for r in (0, 1):                               # one rank re-configures to the SAME size: only it asks
            ctxs[r].call("lg_image_configure", 640, 360)
            d = frame([(640, 360)] * 2)
            assert (np.abs(d - c_) <= 1e-5 * np.maximum(1.0, np.abs(c_))).all()
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.skipif(not have_cuda() or device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", [0, NCCL, PEER], ids=["auto", "nccl", "peer"])
def test_reduce_of_images_of_different_sizes_is_an_error_on_every_rank(oracle, mode):
    """A rank whose frame has another size must not reach ncclReduce (different counts per rank hang or corrupt) nor the
    peer kernel: every rank gets LG_ERR_INVALID, and the next reduce with equal sizes and a root other than 0 works."""
    from light_garden_b200 import _lib
    from light_garden_b200._lib import LightGardenError
    from light_garden_b200.tracer import Context, Renderer, Tracer
    spec = small_specs()["C1"]
    ctxs = [Context(d, abi.LG_PRECISION_F32) for d in range(2)]
    try:
        arr = (C.c_void_p * 2)(*[c.h for c in ctxs])
        _lib.check(ctxs[0].h, _lib.load().lg_comm_init_all(arr, 2))
        tracers = []
        for rk, c in enumerate(ctxs):
            c.call("lg_reduce_mode_set", mode)
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=c))
            t.set_shard(rk, 2)
            tracers.append(t)

        def reduce_collect(root):
            errs = [None, None]

            def run(i):
                try:
                    ctxs[i].call("lg_image_reduce", root, C.byref(C.c_float()))
                except LightGardenError as e:
                    errs[i] = e
            th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
            [x.start() for x in th]
            [x.join(120) for x in th]
            assert not any(x.is_alive() for x in th), "lg_image_reduce hangs"
            return errs

        rends = [Renderer(ctxs[0], 480, 270), Renderer(ctxs[1], 640, 360)]
        for rk, r in enumerate(rends):
            r.clear(1.0 if rk == 0 else 0.0)
            r.render(tracers[rk])
        errs = reduce_collect(0)
        assert all(e is not None and e.code == abi.LG_ERR_INVALID and "size" in e.message for e in errs), errs
        rends = [Renderer(c, 480, 270) for c in ctxs]
        for rk, r in enumerate(rends):
            r.clear(1.0 if rk == 1 else 0.0)                 # root 1 owns the clear alpha this time
            r.render(tracers[rk])
        parts = [r.read_rgba32f() for r in rends]
        assert reduce_collect(1) == [None, None]
        total = rends[1].read_rgba32f()
        assert (np.abs(total - (parts[0] + parts[1])) <= 1e-6 * np.maximum(1.0, np.abs(total))).all()
    finally:
        for c in ctxs:
            c.close()
