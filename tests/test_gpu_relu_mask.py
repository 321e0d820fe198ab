"""cb_gemm_rows_masked: the forward transform that also writes the relu gate as bytes (the input Linear + relu,
GCN.py:104-106, whose backward rides on the first layer's adjoint GEMM)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('M,K,N', [(1000, 64, 64), (4097, 128, 256), (130, 256, 128)])
def test_masked_transform_equals_plain_transform_and_its_sign(dtype, M, K, N):
    from gnn_tail_generalization_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).to(DEV).to(dtype)
    W = (torch.randn(K, N, generator=g) * 0.1).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    wt = ops.split_weight(W, transpose=True, dtype=dtype)
    plain = ops.gemm_rows_raw(A, wt, bias=bias, relu=True)
    out, mask = ops.gemm_rows_raw(A, wt, bias=bias, relu=True, want_relu_mask=True)
    assert torch.equal(out, plain)
    assert mask.dtype == torch.uint8 and torch.equal(mask.bool(), out > 0)
    # the gradient GEMM gated by the bytes == gated by the activations
    dy = torch.randn(M, N, generator=g).to(DEV).to(dtype)
    wb = ops.split_weight(torch.randn(N, N, generator=g).to(DEV) * 0.1, transpose=False, dtype=dtype)
    a, ca, _ = ops.gemm_rows_grad_raw(dy, wb, gate_u8=mask, want_col_sum=True)
    b, cb, _ = ops.gemm_rows_grad_raw(dy, wb, gate_f32=out, want_col_sum=True)
    assert torch.equal(a, b) and torch.equal(ca, cb)
