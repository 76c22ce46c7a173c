"""Bitwise run-to-run check of the hot-path ops (same process, same inputs, fresh L2 state in between): every output and
gradient of two passes must be identical -- the library has no atomics on floating-point data in these paths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adaptsegnet_b200 import ops
from adaptsegnet_b200.model.deeplab_multi import Classifier_Module
from adaptsegnet_b200.model.discriminator import FCDiscriminator

dev = "cuda"
torch.manual_seed(3)
ok = True
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for (HW, hw, cl) in (((256, 128), (33, 17), False), ((720, 1280), (90, 160), True), ((512, 1024), (65, 129), True)):
    head = Classifier_Module(2048, [6, 12, 18, 24], [6, 12, 18, 24], 19).to(dev)
    D = FCDiscriminator(19).to(dev)
    f = (torch.randn(1, 2048, *hw, device=dev).abs() * 1.6)
    if cl:
        f = f.contiguous(memory_format=torch.channels_last)
    lab = torch.randint(0, 19, (1,) + HW, device=dev)
    outs = []
    for rep in range(2):
        flush.zero_()
        head._pack.invalidate(); D._pack.invalidate()
        for p in list(head.parameters()) + list(D.parameters()):
            p.grad = None
        x = f.clone().requires_grad_(True)
        z = head(x)
        loss = ops.upsample_softmax_cross_entropy(z, HW, lab)
        d = D(z, from_logits=True, up_size=HW)
        l2 = ops.gan_loss(d, 0.0, ops.GAN_BCE)
        (loss + l2).backward()
        outs.append([z.detach().clone(), loss.detach().clone(), d.detach().clone(), x.grad.clone()] +
                    [p.grad.clone() for p in list(head.parameters()) + list(D.parameters())])
    same = [torch.equal(a, b) for a, b in zip(*outs)]
    print(HW, hw, "channels_last" if cl else "nchw", "bitwise identical:", all(same), [i for i, s in enumerate(same) if not s])
    ok = ok and all(same)
print("DETERMINISTIC" if ok else "NOT DETERMINISTIC")
sys.exit(0 if ok else 1)
