import torch, sys
sys.path.insert(0, '.')
from townlet_b200 import capi
from townlet_b200.params import ObsSpec
from townlet_b200.rollout import FusedPolicyWeights, PolicyMLP, policy_step
obs = ObsSpec()
torch.manual_seed(0)
policy = PolicyMLP(obs.n_features, obs.map_shape, action_dim=9).cuda()
w = FusedPolicyWeights(policy, torch.device("cuda"))
n = 65536
x = torch.zeros((n, 640), dtype=torch.bfloat16, device="cuda")
x[:, :484] = (torch.rand((n, 484), device="cuda") < 0.05).to(torch.bfloat16)
x[:, 512:593] = torch.randn((n, 81), device="cuda").to(torch.bfloat16)
out = (torch.empty(n, dtype=torch.int64, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, device="cuda"))
for pair in (0, 1, 2):  # single CTAs, CTA pairs (cta_group::2), multicast clusters
    capi.set_option("policy_pair", 1 if pair == 1 else 0)
    capi.set_option("policy_multicast", 1 if pair == 2 else 0)
    print("form", pair, flush=True)
    for _ in range(3):
        policy_step(x, w, seed=1, out=out)
        torch.cuda.synchronize()
