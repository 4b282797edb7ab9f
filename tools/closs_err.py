import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from dg_tta_b200.tta.torch_utils import consistency_dice_loss
from test_consistency_gpu import reference_loss
torch.manual_seed(0)
for scale in (1.0, 3.0, 8.0):
    a = (torch.randn(2, 14, 64, 64, 64, device="cuda") * scale + 0.3)
    b = (torch.randn(2, 14, 64, 64, 64, device="cuda") * scale + 0.3)
    a1 = a.clone().double().requires_grad_(True)
    ref = reference_loss(a1, b.double()); ref.backward()
    a2 = a.clone().requires_grad_(True)
    got = consistency_dice_loss(a2, b); got.backward()
    a3 = a.clone().requires_grad_(True)
    t32 = reference_loss(a3, b); t32.backward()
    g = a1.grad.abs().max().item()
    print(f"scale {scale}: loss err ours {abs(got.item()-ref.item()):.2e} torch32 {abs(t32.item()-ref.item()):.2e}; grad relerr ours {(a2.grad.double()-a1.grad).abs().max().item()/g:.2e} torch32 {(a3.grad.double()-a1.grad).abs().max().item()/g:.2e}")
