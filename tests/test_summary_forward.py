"""The var-len summary-forward adapter (cogreasoner_chat.py:306-316): its orchestration (restarting position ids,
layer loop over the frozen HF modules) against ``Qwen2Model.forward`` itself, one sequence at a time as the reference calls it."""
import pytest
import torch

from cogstream_b200.summary_forward import make_varlen_summary_forward, packed_position_ids


def _tiny_qwen2(dtype=torch.float32, device="cpu"):
    from transformers import Qwen2Config, Qwen2Model
    torch.manual_seed(0)
    cfg = Qwen2Config(vocab_size=512, hidden_size=64, intermediate_size=128, num_hidden_layers=3, num_attention_heads=4,
                      num_key_value_heads=2, max_position_embeddings=512)
    return Qwen2Model(cfg).to(device=device, dtype=dtype).eval()


def test_packed_position_ids():
    cu = torch.tensor([0, 3, 3, 7], dtype=torch.int32)
    assert packed_position_ids(cu, 7).tolist() == [0, 1, 2, 0, 1, 2, 3]


def test_varlen_forward_equals_sequential_hf_forwards_cpu():
    """fp32 on the CPU with the per-sequence SDPA loop: the packed pass reproduces the reference's k batch-1 forwards."""
    llm = _tiny_qwen2()
    lens = [5, 1, 17, 9]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    x = torch.randn(sum(lens), 64)
    fwd = make_varlen_summary_forward(llm, attention="sdpa_loop")
    got = fwd(x, cu)
    b = cu.tolist()
    with torch.no_grad():
        for j in range(len(lens)):
            e = x[b[j]: b[j + 1]].unsqueeze(0)
            want = llm(position_ids=None, attention_mask=torch.ones(1, lens[j], dtype=torch.long), inputs_embeds=e).last_hidden_state[0]
            assert torch.allclose(got[b[j]: b[j + 1]], want, atol=2e-5, rtol=1e-4), j


@pytest.mark.gpu
@pytest.mark.parametrize("attention", ["flash_attn", "torch_varlen"])
def test_varlen_forward_on_gpu_feeds_fused_selection(attention):
    """bf16 on the GPU with a real var-len attention kernel: hidden states within bf16 kernel noise of the sequential HF
    forwards, and the packed hidden states go through csb_relevance_select's fused mean pool bit-identically to
    csb_seq_mean_pool."""
    from cogstream_b200 import ops
    dev = torch.device("cuda:0")
    llm = _tiny_qwen2(torch.bfloat16, dev)
    try:
        fwd = make_varlen_summary_forward(llm, attention=attention)
    except Exception as exc:
        pytest.skip(f"{attention} not importable: {exc}")
    lens = [33, 7, 64, 21, 50]
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device=dev)
    x = torch.randn(sum(lens), 64, device=dev).to(torch.bfloat16)
    try:
        got = fwd(x, cu)
    except Exception as exc:  # a library kernel that does not run on this device is not our failure
        pytest.skip(f"{attention} does not run here: {type(exc).__name__}: {exc}")
    b = cu.tolist()
    with torch.no_grad():
        for j in range(len(lens)):
            e = x[b[j]: b[j + 1]].unsqueeze(0)
            want = llm(position_ids=None, attention_mask=torch.ones(1, lens[j], dtype=torch.long, device=dev), inputs_embeds=e).last_hidden_state[0]
            err = float((got[b[j]: b[j + 1]].float() - want.float()).abs().max() / want.float().abs().max())
            assert err < 4e-2, (j, err)
    pooled = ops.seq_mean_pool(got.contiguous(), cu)
    ref = torch.stack([got[b[j]: b[j + 1]].float().mean(0) for j in range(len(lens))])
    assert float((pooled.float() - ref).abs().max()) < 2e-2
