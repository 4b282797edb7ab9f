import sys, torch
sys.path.insert(0, '/root/repo')
from dg_tta_b200.tta.torch_utils import consistency_dice_loss
ta = torch.randn(2, 14, 128, 128, 128, device="cuda", requires_grad=True)
tb = torch.randn(2, 14, 128, 128, 128, device="cuda")
for _ in range(3):
    torch.autograd.grad(consistency_dice_loss(ta, tb), ta)
torch.cuda.synchronize()
