"""CPU: checkpoint loading rules of det3d/torchie/trainer/checkpoint.py:67-137,166-218 (module. prefix, state_dict
wrapper, spconv 1.x -> 2.x weight layout, reporting instead of raising)."""
import logging
import os

import pytest
import torch


def _backbone():
    import pillarnet_lts_b200  # noqa: F401
    from pillarnet_lts_b200.backbone import PillarResNet18S
    torch.manual_seed(0)
    return PillarResNet18S(in_channels=32)


def test_roundtrip_with_module_prefix_and_spconv1_layout(tmp_path):
    from pillarnet_lts_b200.checkpoint import load_checkpoint
    src, dst = _backbone(), _backbone()
    for p in src.parameters():
        torch.nn.init.normal_(p)
    sd = {}
    for k, v in src.state_dict().items():
        if v.dim() == 4 and "conv" in k and k.endswith(".weight"):      # sparse conv: save it the spconv-1.x way
            v = v.permute(1, 2, 3, 0).contiguous()                      # (Cout,kH,kW,Cin) -> (kH,kW,Cin,Cout)
        sd["module." + k] = v
    sd["module.not_in_the_model"] = torch.zeros(3)
    path = os.path.join(tmp_path, "ckpt.pth")
    torch.save({"state_dict": sd, "meta": {"epoch": 20}}, path)
    ck = load_checkpoint(dst, path, map_location="cpu", logger=logging.getLogger("t"))
    assert ck["meta"]["epoch"] == 20
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k


def test_strict_raises_and_missing_files_are_errors(tmp_path):
    from pillarnet_lts_b200.checkpoint import load_checkpoint, load_state_dict
    m = _backbone()
    sd = dict(m.state_dict())
    sd.pop(next(k for k in sd if k.endswith("bias")))
    with pytest.raises(RuntimeError):
        load_state_dict(m, sd, strict=True)
    rep = load_state_dict(m, sd, strict=False, logger=logging.getLogger("t"))
    assert len(rep["missing"]) == 1 and not rep["unexpected"]
    with pytest.raises(IOError):
        load_checkpoint(m, os.path.join(tmp_path, "nope.pth"))


def test_pretrained_kwarg_goes_through_load_checkpoint_and_save_roundtrips(tmp_path):
    """SingleStageDetector(pretrained=...) must use the checkpoint loader (module. prefix from a DDP save,
    spconv-1.x sparse layout, non-tensor objects in `meta`), as detectors/single_stage.py:29-37 does; and
    save_checkpoint writes the reference's {meta, state_dict, optimizer} file."""
    import pillarnet_lts_b200 as P
    from pillarnet_lts_b200 import configs
    from pillarnet_lts_b200.checkpoint import save_checkpoint
    from pillarnet_lts_b200.registry import ConfigDict
    cfg = configs.get("nusc18")
    cfg["model"]["neck"]["layer_nums"] = [1, 1]
    torch.manual_seed(3)
    src = P.build_detector(ConfigDict.wrap(cfg["model"]), None, ConfigDict.wrap(cfg["test_cfg"]))
    for p in src.parameters():
        torch.nn.init.normal_(p, std=0.1)
    sd = {}
    for k, v in src.state_dict().items():
        if v.dim() == 4 and k.startswith("backbone.conv") and "conv5" not in k and k.endswith(".weight"):
            v = v.permute(1, 2, 3, 0).contiguous()          # a spconv-1.x checkpoint
        sd["module." + k] = v
    path = os.path.join(tmp_path, "ddp_1x.pth")

    import argparse                                         # a non-tensor object: rejected by weights_only=True
    torch.save({"state_dict": sd, "meta": {"obj": argparse.Namespace(epoch=7)}}, path)
    cfg2 = configs.get("nusc18")
    cfg2["model"]["neck"]["layer_nums"] = [1, 1]
    cfg2["model"]["pretrained"] = path
    torch.manual_seed(99)
    dst = P.build_detector(ConfigDict.wrap(cfg2["model"]), None, ConfigDict.wrap(cfg2["test_cfg"]))
    for (k, a), (_, b) in zip(src.state_dict().items(), dst.state_dict().items()):
        assert torch.equal(a, b), k
    # save -> load round trip in the reference's file layout
    opt = torch.optim.SGD(dst.parameters(), lr=0.1)
    out = save_checkpoint(dst, os.path.join(tmp_path, "sub", "epoch_1.pth"), optimizer=opt, meta={"epoch": 1})
    ck = torch.load(out, map_location="cpu", weights_only=False)
    assert set(ck) == {"meta", "state_dict", "optimizer"} and ck["meta"]["epoch"] == 1
    assert list(ck["state_dict"].keys()) == list(src.state_dict().keys())
    assert all(not v.is_cuda for v in ck["state_dict"].values())
