"""Attention kernels vs a plain PyTorch fp32 reference of the same op (softmax(q k^T/sqrt(dh) + key mask) v).
impl 0 = fp32 CUDA-core kernel (parity mode): tolerance 2e-2 is the bf16 rounding of the OUTPUT only (values O(1));
impl 1 = tcgen05 kernel: q, k, v and P are bf16 operands -> tolerance 3e-2 abs on O(1) outputs."""
import math

import pytest
import torch

ATTN6_DEFAULT = 5   # csrc/attention_tc6.cuh opt_attn6(): generation-6 kernel with P handed over in two halves

pytestmark = pytest.mark.gpu


def _ref(q, k, v, mask):
    s = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    if mask is not None:
        s = s.masked_fill(~mask[:, None, None, :].bool(), float("-inf"))
    o = s.softmax(-1) @ v
    return o.permute(0, 2, 1, 3).reshape(q.shape[0], q.shape[2], -1)


# 0 = fp32 CUDA-core kernel (parity mode); 1 = the tcgen05 kernel the product uses (generation 6, attention_tc6.cuh, unless the option attn6 says
# otherwise); 4 = generation 4 (attention_tc4.cuh); +100 = q / k rows of 80 elements for dh = 72 (160-byte pitch) instead of 128
@pytest.mark.parametrize("impl", [0, 1, 4, 101, 104])
@pytest.mark.parametrize("B,H,Lq,Lk,dh,masked", [(2, 4, 500, 500, 72, False), (2, 3, 256, 256, 64, False), (3, 2, 500, 100, 72, True),
                                                 (2, 2, 40, 12, 72, True), (1, 2, 130, 130, 64, False), (1, 16, 1500, 1500, 72, False),
                                                 (2, 2, 37, 100, 64, True), (8, 16, 500, 500, 72, False), (2, 5, 700, 700, 72, "grow")])
def test_attention(impl, B, H, Lq, Lk, dh, masked):
    from ezaudio_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(Lq * 7 + Lk + dh)
    q = torch.randn(B, H, Lq, dh, device="cuda", generator=g) * 1.5
    k = torch.randn(B, H, Lk, dh, device="cuda", generator=g) * 1.5
    v = torch.randn(B, H, Lk, dh, device="cuda", generator=g)
    if masked == "grow":  # scores grow by orders of magnitude from key block to key block: exercises the in-place O rescale
        k = k * torch.linspace(0.2, 3.0, Lk, device="cuda")[None, None, :, None]
        masked = False
    mask = None
    if masked:
        mask = torch.zeros(B, Lk, dtype=torch.uint8, device="cuda")
        for i in range(B):
            mask[i, : (1 if i == B - 1 else min(Lk, 8 + 5 * i))] = 1
    out = torch.zeros(B, Lq, H * dh, device="cuda", dtype=torch.bfloat16)
    L = _lib.lib()
    if impl == 0:
        args = (q.contiguous(), k.contiguous(), v.contiguous())
        ref = _ref(q, k, v, mask)
    else:
        dhp, dvp, lkp = (dh + 63) // 64 * 64, (dh + 15) // 16 * 16, (Lk + 7) // 8 * 8
        if impl >= 100 and dh == 72:
            dhp = 80
        qb = torch.zeros(B * H, Lq, dhp, device="cuda", dtype=torch.bfloat16)
        kb = torch.zeros(B * H, Lk, dhp, device="cuda", dtype=torch.bfloat16)
        vt = torch.zeros(B * H, dvp, lkp, device="cuda", dtype=torch.bfloat16)
        qb[:, :, :dh] = q.reshape(B * H, Lq, dh)
        kb[:, :, :dh] = k.reshape(B * H, Lk, dh)
        vt[:, :dh, :Lk] = v.reshape(B * H, Lk, dh).transpose(1, 2)
        vt[:, :, Lk:] = 7.0  # beyond the true length: must never be read
        args = (qb, kb, vt)
        ref = _ref(q.bfloat16().float(), k.bfloat16().float(), v.bfloat16().float(), mask)
    _run(L, _lib, args, mask, out, B, H, Lq, Lk, dh, impl)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert math.isfinite(err) and err < (2e-2 if impl == 0 else 3e-2), err


def _run(L, _lib, args, mask, out, B, H, Lq, Lk, dh, impl):
    _lib.check(L.ezb_test_attention(0, _lib.ptr(args[0]), _lib.ptr(args[1]), _lib.ptr(args[2]), _lib.ptr(mask), _lib.ptr(out), B, H, Lq, Lk, dh, impl,
                                    _lib.stream_ptr()))
    torch.cuda.synchronize()


@pytest.mark.parametrize("B,H,Lq,Lk,dh,masked", [(2, 4, 500, 500, 72, False), (3, 2, 500, 100, 72, True), (2, 2, 40, 12, 72, True), (1, 16, 1500, 1500, 72, False),
                                                 (2, 3, 256, 256, 64, False), (8, 16, 500, 500, 72, False), (2, 5, 400, 512, 72, "grow"), (5, 3, 300, 100, 64, True),
                                                 (16, 16, 500, 500, 72, False), (2, 2, 130, 385, 72, True)])
def test_attention_kv_resident(B, H, Lq, Lk, dh, masked):
    """attn4 with the K / V^T key blocks of a head resident in shared memory (option attn_res; falls back above 512 keys): odd and even numbers
    of query tiles per head, one to four key blocks, more heads than SMs (two waves of CTAs), masks, growth."""
    from ezaudio_b200 import _lib
    L = _lib.lib()
    _lib.check(L.ezb_set_option(b"attn6", 0))   # generation-4 kernel (the default is generation 6)
    _lib.check(L.ezb_set_option(b"attn_res", 1))
    try:
        test_attention(1, B, H, Lq, Lk, dh, masked)
        test_attention(101, B, H, Lq, Lk, dh, masked)
    finally:
        _lib.check(L.ezb_set_option(b"attn_res", 0))
        _lib.check(L.ezb_set_option(b"attn6", ATTN6_DEFAULT))


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,Lq,Lk,dh,masked", [(2, 4, 500, 500, 72, False), (3, 2, 500, 100, 72, True), (2, 2, 40, 12, 72, True), (1, 16, 1500, 1500, 72, False),
                                                 (2, 3, 256, 256, 64, False), (8, 16, 500, 500, 72, False), (2, 5, 400, 512, 72, "grow"), (5, 3, 300, 100, 64, True),
                                                 (16, 16, 500, 500, 72, False), (2, 2, 130, 385, 72, True), (1, 1, 100, 300, 72, False), (1, 3, 128, 128, 72, True)])
@pytest.mark.parametrize("res", [0, 1])
def test_attention_mufu_token(B, H, Lq, Lk, dh, masked, res):
    """attn4 with the exponent phases of the two softmax groups strictly alternating (option attn_pp): CTAs with an odd and an even number of
    items (the group without a last item keeps passing the token), a single item, one to twelve key blocks, with and without resident K / V^T."""
    from ezaudio_b200 import _lib
    L = _lib.lib()
    _lib.check(L.ezb_set_option(b"attn6", 0))   # generation-4 kernel (the default is generation 6)
    _lib.check(L.ezb_set_option(b"attn_pp", 1))
    _lib.check(L.ezb_set_option(b"attn_res", res))
    try:
        test_attention(1, B, H, Lq, Lk, dh, masked)
        test_attention(101, B, H, Lq, Lk, dh, masked)
    finally:
        _lib.check(L.ezb_set_option(b"attn_pp", 0))
        _lib.check(L.ezb_set_option(b"attn_res", 0))
        _lib.check(L.ezb_set_option(b"attn6", ATTN6_DEFAULT))


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,Lq,Lk,dh,masked", [(2, 4, 500, 500, 72, False), (3, 2, 500, 100, 72, True), (2, 2, 40, 12, 72, True), (1, 16, 1500, 1500, 72, False),
                                                 (2, 3, 256, 256, 64, False), (8, 16, 500, 500, 72, False), (2, 5, 400, 512, 72, "grow"), (5, 3, 300, 100, 64, True),
                                                 (16, 16, 500, 500, 72, False), (2, 2, 130, 385, 72, True), (1, 1, 100, 300, 72, False), (1, 3, 128, 128, 72, True)])
@pytest.mark.parametrize("mode", [0, 1, 3, 5, 7])
def test_attention_gen6(B, H, Lq, Lk, dh, masked, mode):
    """attn6 (attention_tc6.cuh: chunked two-pass softmax, packed f32x2 arithmetic; bit 1 of the option = MUFU token between the two softmax
    groups, bit 2 = P handed to the MMA warp in two halves): every mode on odd / even item counts per CTA, a single item, one to twelve key
    blocks, key masks, score growth (in-place O rescale before the exponent phase), dh = 64 and 72, both q / k row pitches."""
    from ezaudio_b200 import _lib
    L = _lib.lib()
    _lib.check(L.ezb_set_option(b"attn6", mode))
    try:
        test_attention(1, B, H, Lq, Lk, dh, masked)
        test_attention(101, B, H, Lq, Lk, dh, masked)
    finally:
        _lib.check(L.ezb_set_option(b"attn6", ATTN6_DEFAULT))


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,Lq,Lk,dh,masked", [(2, 4, 500, 500, 72, False), (3, 2, 500, 100, 72, True), (2, 2, 40, 12, 72, True), (1, 16, 1500, 1500, 72, False),
                                                 (2, 3, 256, 256, 64, False), (8, 16, 500, 500, 72, False), (2, 5, 400, 512, 72, "grow"), (5, 3, 300, 100, 64, True),
                                                 (16, 16, 500, 500, 72, False), (2, 2, 130, 385, 72, True), (1, 1, 100, 300, 72, False), (1, 3, 128, 128, 72, True),
                                                 (3, 5, 200, 65, 72, True), (1, 2, 640, 192, 64, False), (37, 4, 512, 512, 72, False)])
def test_attention_gen7(B, H, Lq, Lk, dh, masked):
    """attn7 (attention_tc7.cuh: 64-key score blocks, two S buffers per group, scores issued two blocks ahead, output stores issued by the Q producer
    warp): odd / even item counts per CTA, a single item, an odd number of 64-key blocks per item (385, 65, 192 keys), a lone key in the last block,
    up to 24 blocks, key masks, score growth (in-place O rescale), dh = 64 and 72, both q / k row pitches, Lk <= 64 (falls back to generation 6),
    exactly four items on every CTA."""
    test_attention(7, B, H, Lq, Lk, dh, masked)
    test_attention(107, B, H, Lq, Lk, dh, masked)
