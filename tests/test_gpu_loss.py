"""-m gpu: the fused TotalLoss against fixtures made by the reference's TotalLoss + autograd."""
import pytest
import torch

from tests import cases

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.mark.parametrize("name", list(cases.LOSS_CASES))
def test_total_loss_and_gradients_match_reference_golden(golden, name):
    from object_nerf_b200.losses import TotalLoss
    maps, batch = cases.build_loss_case(cases.LOSS_CASES[name])
    gold = golden(name)
    maps = {k: v.to(DEV).requires_grad_(True) for k, v in maps.items()}
    batch = {k: v.to(DEV) for k, v in batch.items()}
    loss_sum, loss_dict = TotalLoss(dict(cases.LOSS_CONF))(maps, batch)
    # masked means: fp64 accumulation here vs torch's fp32 sum -> 1e-6 relative
    assert abs(loss_sum.item() - gold["loss_sum"].item()) <= 2e-6 * abs(gold["loss_sum"].item())
    assert {"term|" + k for k in loss_dict} == {k for k in gold if k.startswith("term|")}    # skipped terms are absent
    for k, v in loss_dict.items():
        assert abs(v.item() - gold["term|" + k].item()) <= 2e-6 * abs(gold["term|" + k].item()) + 1e-12, k
    loss_sum.backward()
    for k, v in maps.items():
        want = gold["grad|" + k]
        got = v.grad.cpu() if v.grad is not None else torch.zeros_like(want)
        assert ((got - want).abs() <= 3e-6 * want.abs() + 1e-10).all(), (k, (got - want).abs().max().item())
        assert ((want == 0) == (got == 0)).all(), k      # the masks (and the clamp) select exactly the same entries
