"""-m gpu: regression tests for the round-1 advisor findings and the judge's small items."""

import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    os.environ["MOJO_BACKEND"] = "b200"
    import mojo_opset_b200 as m

    return m


@pytest.mark.parametrize("max_q_len", [64, 400], ids=["general-kernel", "tcgen05-kernel"])
def test_global_only_window_of_zero_reads_as_zeros(ops, max_q_len):
    """global_window_size=0, local None: no key is visible to any row.  Both attention kernels must leave zeros, not
    the uninitialised contents of the freshly allocated output."""
    Hq, Hkv, D, bs, T = 4, 2, 128, 16, max_q_len
    nb = T // bs + 2
    kc = torch.randn(nb, Hkv, bs, D, device=DEV).to(torch.bfloat16)
    vc = torch.randn(nb, Hkv, bs, D, device=DEV).to(torch.bfloat16)
    q = torch.randn(T, Hq, D, device=DEV).to(torch.bfloat16)
    table = torch.arange(nb, dtype=torch.int32, device=DEV).view(1, -1)
    cu = torch.tensor([0, T], dtype=torch.int32, device=DEV)
    # dirty the allocator's free blocks so that an unwritten output would not read as zeros by luck
    junk = torch.full((T, Hq, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    del junk
    op = ops.MojoPagedPrefillSWA(global_window_size=0, local_window_size=None)
    out = op(q, kc, vc, cu, table, max_q_len=max_q_len)
    assert out.shape == q.shape and not out.float().abs().sum().item()


def test_rotary_embedding_survives_model_to_bf16(ops):
    """``model.to(torch.bfloat16)`` casts the op's ``inv_freq`` / table buffers; the b200 op must keep working."""
    rot = ops.MojoRotaryEmbedding(1e6, 128, init_max_length=64, device=DEV)
    x = torch.empty(7, 4096, device=DEV, dtype=torch.bfloat16)
    pos = torch.tensor([0, 1, 5, 9, 33, 62, 63], dtype=torch.int32, device=DEV)
    cos32, sin32 = rot(x, position_ids=pos)
    rot = rot.to(torch.bfloat16)
    assert rot.inv_freq.dtype == torch.bfloat16
    cos16, sin16 = rot(x, position_ids=pos)
    torch.testing.assert_close(cos16.float(), cos32.float(), atol=2e-2, rtol=2e-2)
    torch.testing.assert_close(sin16.float(), sin32.float(), atol=2e-2, rtol=2e-2)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ops_run_on_the_tensors_device_not_the_current_one(ops):
    """Tensors on cuda:1 while cuda:0 is current (HF device_map / accelerate placement)."""
    from oracle import golden

    assert torch.cuda.current_device() == 0
    dev = "cuda:1"
    g = torch.Generator().manual_seed(3)
    x, r = (torch.randn(9, 1024, generator=g).to(torch.bfloat16) for _ in range(2))
    w = torch.randn(1024, generator=g).to(torch.bfloat16)
    norm = ops.MojoResidualAddRMSNorm(1024, eps=1e-6, device=dev, dtype=torch.bfloat16)
    with torch.no_grad():
        norm.weight.copy_(w)
    y, s = norm(x.to(dev), r.to(dev))
    y_ref, s_ref = golden.residual_add_rms_norm(x, r, w, 1e-6)
    assert y.device == torch.device(dev) and torch.equal(s.cpu(), s_ref)
    torch.testing.assert_close(y.cpu().float(), y_ref.float(), atol=5e-2, rtol=1e-2)
    B, Hq, Hkv, D, bs, ctx = 3, 8, 2, 128, 16, 300
    nb = B * 20
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(torch.bfloat16)
    table = torch.randperm(nb, generator=g)[: B * 19].view(B, 19).to(torch.int32)
    lens = torch.tensor([ctx, 17, 250], dtype=torch.int32)
    q = torch.randn(B, Hq, D, generator=g).to(torch.bfloat16)
    out = ops.MojoPagedDecodeGQA()(q.to(dev), kc.to(dev), vc.to(dev), lens.to(dev), table.to(dev))
    ref = golden.paged_decode_gqa(q, kc, vc, lens, table)
    torch.testing.assert_close(out.cpu().float(), ref.float(), atol=2e-2, rtol=2e-2)
    assert torch.cuda.current_device() == 0


def test_unmapped_first_block_raises_value_error_like_the_reference(ops):
    """Reference attention.py:186-187 / 396-397: a row with keys whose first block id is negative is a ValueError.  A host
    check would synchronise; the kernels flag it in a device word that ``check_device_errors()`` turns into the error."""
    import mojo_opset_b200 as m

    B, Hq, Hkv, D, bs = 3, 8, 2, 128, 16
    kc = torch.randn(12, Hkv, bs, D, device=DEV).to(torch.bfloat16)
    vc = torch.randn(12, Hkv, bs, D, device=DEV).to(torch.bfloat16)
    q = torch.randn(B, Hq, D, device=DEV).to(torch.bfloat16)
    table = torch.tensor([[0, 1, 2], [-1, 4, 5], [6, 7, 8]], dtype=torch.int32, device=DEV)
    lens = torch.tensor([40, 20, 0], dtype=torch.int32, device=DEV)
    try:
        m.check_device_errors()                  # clean slate (earlier tests may have fed unmapped tables on purpose)
    except ValueError:
        pass
    ops.MojoPagedDecodeGQA()(q, kc, vc, lens, table)
    with pytest.raises(ValueError, match="Paged decode requires a valid block table"):
        m.check_device_errors()
    m.check_device_errors()                      # the word was cleared
    lens_ok = torch.tensor([40, 0, 0], dtype=torch.int32, device=DEV)   # the unmapped row has no keys: fine
    ops.MojoPagedDecodeGQA()(q, kc, vc, lens_ok, table)
    m.check_device_errors()
    qp = torch.randn(30, Hq, D, device=DEV).to(torch.bfloat16)
    cu = torch.tensor([0, 10, 30, 30], dtype=torch.int32, device=DEV)
    ops.MojoPagedPrefillGQA()(qp, kc, vc, cu, table, max_q_len=20)
    with pytest.raises(ValueError, match="Paged prefill requires a valid block table"):
        m.check_device_errors()


@pytest.mark.parametrize("head_dim,dtype", [(128, torch.bfloat16), (64, torch.float16)])
@pytest.mark.parametrize("splits", [2, 5, 37])
def test_split_kv_fold_in_kernel_matches_fold_kernel_and_oracle(ops, splits, head_dim, dtype, monkeypatch):
    """Split-KV decode: the last split of a (sequence, kv head) group to arrive folds the group's partials inside the
    main kernel (arrival counters registered with `mojo_b200_set_decode_tickets`).  Both fold paths must agree with the
    oracle on a ragged batch with empty sequences and empty splits, over repeated launches (the counters must be back
    at zero after every launch) and for more rows than one 16-head tile."""
    from mojo_opset_b200 import _lib
    from oracle import golden

    g = torch.Generator().manual_seed(splits)
    Hq, Hkv, D, bs = 40, 2, head_dim, 16  # group 20: two head tiles per kv head
    lens = [2500, 0, 70, 1, 1023, 64, 2400]
    B, max_len = len(lens), 2560
    nblk = max_len // bs
    nb = B * nblk + 3
    kc = torch.randn(nb, Hkv, bs, D, generator=g).to(dtype).to(DEV)
    vc = torch.randn(nb, Hkv, bs, D, generator=g).to(dtype).to(DEV)
    q = torch.randn(B, Hq, D, generator=g).to(dtype).to(DEV)
    table = torch.full((B, nblk), -1, dtype=torch.int32)
    perm = torch.randperm(nb, generator=g)
    at = 0
    for i, n in enumerate(lens):
        need = (n + bs - 1) // bs
        table[i, :need] = perm[at:at + need].to(torch.int32)
        at += need
    table, seq = table.to(DEV), torch.tensor(lens, dtype=torch.int32, device=DEV)
    ref = golden.paged_decode_gqa(q, kc, vc, seq, table)
    op = ops.MojoPagedDecodeGQA()
    monkeypatch.setenv("MOJO_B200_DECODE_SPLITS", str(splits))
    outs = {}
    for fold in ("1", "0"):
        monkeypatch.setenv("MOJO_B200_DECODE_FOLD", fold)
        for _ in range(3):
            outs[fold] = op(q, kc, vc, seq, table, max_total_seq_len=max_len)
        torch.testing.assert_close(outs[fold].float(), ref.float(), atol=2e-2, rtol=2e-2)
        assert not outs[fold][1].float().abs().sum().item()  # the empty sequence reads as zeros
    torch.testing.assert_close(outs["1"].float(), outs["0"].float(), atol=4e-3, rtol=1e-2)  # summation order only
    torch.cuda.synchronize()
    tickets = _lib._decode_tickets[torch.cuda.current_device()]
    assert int(tickets.abs().sum().item()) == 0, "arrival counters must be zero between launches"
