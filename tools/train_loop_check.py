import sys, torch
sys.path.insert(0, '/root/repo')
import bench
from splatco_b200.gaussian_renderer import prefilter_voxel, render
from splatco_b200.loss import l1_ssim_loss, scaling_reg
dev = torch.device("cuda", 0)
cfg = bench.WORKLOADS["c2"]
pc = bench.build_model(cfg, dev); pc.feat_planes.Q0 = 0.03
cams, gts = bench.build_views(cfg); cams = [c.to(dev) for c in cams]; gts = [g.to(dev) for g in gts]
bg = torch.ones(3, device=dev)
params = [p for p in pc.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-4)
for it in range(61):
    opt.zero_grad(set_to_none=True)
    total = None
    for v in range(cfg["mv"]):
        vm = prefilter_voxel(cams[v], pc, bench.PIPE, bg)
        pkg = render(cams[v], pc, bench.PIPE, bg, visible_mask=vm, retain_grad=True)
        loss = l1_ssim_loss(pkg["render"], gts[v], 0.2) + 0.01 * scaling_reg(pkg["scaling"])
        total = loss if total is None else total + loss
    total.backward()
    opt.step()
    if it in (5, 20, 60):
        torch.cuda.synchronize()
        print(it, f"loss {total.item():.5f} alloc {torch.cuda.memory_allocated()/1e6:.0f} MB reserved {torch.cuda.memory_reserved()/1e6:.0f} MB", flush=True)
