"""GPU parity of the one-pass cross-entropy (csrc/hs_cross_entropy.cu) against F.cross_entropy: loss and gradient,
uint8 and int64 class ids, ignored pixels, upstream gradient scaling; tolerance 2e-6 / 1e-5 (fp32, __expf / __logf)."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,K,P,dtype", [(2, 10, 5000, torch.int64), (3, 1, 777, torch.int64), (1, 32, 4099, torch.uint8),
                                         (8, 10, 98304, torch.uint8), (2, 4, 64, torch.int64)])
def test_cross_entropy_matches_torch(B, K, P, dtype):
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(B * 100 + K)
    logits = (3.0 * torch.randn(B, K, P, generator=g)).to(dev).requires_grad_(True)
    t = torch.randint(0, K, (B, P), generator=g)
    if dtype == torch.int64:
        t[:, ::7] = -100  # ignored pixels (nn.CrossEntropyLoss default ignore_index)
    t = t.to(dtype).to(dev)
    loss = ops.cross_entropy(logits, t)
    (loss * 1.7).backward()
    got_l, got_g = loss.detach().clone(), logits.grad.clone()
    logits.grad = None
    ref = F.cross_entropy(logits, t.long())
    (ref * 1.7).backward()
    assert abs(float(got_l) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
    assert rel_err(got_g.cpu(), logits.grad.cpu()) < 1e-5
    assert ops.STATS.by_name.get("cross_entropy", 0) >= 1


def test_cross_entropy_module_and_fallback():
    from heal_swin_b200 import ops

    dev = torch.device("cuda:0")
    logits = torch.randn(2, 40, 100, device=dev)  # 40 classes: outside the kernel, torch path
    t = torch.randint(0, 40, (2, 100), device=dev)
    assert abs(float(ops.CrossEntropyLoss()(logits, t)) - float(F.cross_entropy(logits, t))) < 1e-6
    logits = torch.randn(2, 5, 8, 9, device=dev)  # 2-D spatial output of the flat twin
    t = torch.randint(0, 5, (2, 8, 9), device=dev)
    assert abs(float(ops.CrossEntropyLoss()(logits, t)) - float(F.cross_entropy(logits, t))) < 1e-6
