"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference from /root/reference (read-only, exists only in the build
container, never on the GPU box) so that golden vectors can be generated from the reference's own code.

Used only by oracle/make_golden.py (and optionally by bench.py's reference arm when a copy of the reference is present
under baseline/_ref). Nothing in the product package imports this module.

The reference imports four third-party modules that are not installed here and are not on the hot path
(SURVEY.md §8c): timm (DropPath / trunc_normal_), mmcv (logger / checkpoint loader), h5py, cosine_annealing_warmup.
They are replaced by minimal stand-ins; BERT weights cannot be downloaded, so `BertModel.from_pretrained` is patched to
build the bert-base-uncased architecture with random weights.
"""
import contextlib
import os
import sys
import tempfile
import types

REFERENCE_CANDIDATES = ["/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "baseline", "_ref")]


def find_reference():
    for p in REFERENCE_CANDIDATES:
        if os.path.exists(os.path.join(p, "lrce", "models", "e2e.py")):
            return os.path.abspath(p)
    return None


def _install_stubs():
    import transformers  # noqa: F401  (must be imported before a spec-less `timm` stub exists)
    import torch

    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")

        class DropPath(torch.nn.Module):
            """stochastic depth; identity outside training (all this harness ever runs)"""

            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                if not self.training or self.drop_prob == 0.0:
                    return x
                keep = 1.0 - self.drop_prob
                mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
                return x * mask / keep

        layers.DropPath = DropPath
        layers.trunc_normal_ = torch.nn.init.trunc_normal_
        timm.models, models.layers = models, layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    if "mmcv" not in sys.modules:
        import logging

        mmcv = types.ModuleType("mmcv")
        utils, runner = types.ModuleType("mmcv.utils"), types.ModuleType("mmcv.runner")
        utils.get_logger = lambda name, log_file=None, log_level=logging.INFO: logging.getLogger(name)
        runner.load_checkpoint = lambda *a, **k: None
        mmcv.utils, mmcv.runner = utils, runner
        sys.modules.update({"mmcv": mmcv, "mmcv.utils": utils, "mmcv.runner": runner})
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    if "cosine_annealing_warmup" not in sys.modules:
        import math

        class CosineAnnealingWarmupRestarts(torch.optim.lr_scheduler.LRScheduler):
            """stand-in for the un-vendored pip-from-git dependency (readme.md:11, agent_base.py:5,56-64): linear warm-up to
            max_lr, cosine decay to min_lr, restart every first_cycle_steps * cycle_mult^k steps with max_lr *= gamma.
            Every param group follows the same schedule, as in the published implementation."""

            def __init__(self, optimizer, first_cycle_steps, cycle_mult=1.0, max_lr=0.1, min_lr=0.001, warmup_steps=0,
                         gamma=1.0, last_epoch=-1):
                self.first, self.mult, self.base_max, self.max_lr = first_cycle_steps, cycle_mult, max_lr, max_lr
                self.min_lr, self.warm, self.gamma = min_lr, warmup_steps, gamma
                self.cur, self.cycle, self.step_in_cycle = first_cycle_steps, 0, last_epoch
                super().__init__(optimizer, last_epoch)
                for g in self.optimizer.param_groups:
                    g["lr"] = self.min_lr

            def get_lr(self):
                if self.step_in_cycle < 0:
                    return [self.min_lr for _ in self.optimizer.param_groups]
                if self.step_in_cycle < self.warm:
                    lr = (self.max_lr - self.min_lr) * self.step_in_cycle / max(self.warm, 1) + self.min_lr
                else:
                    t = (self.step_in_cycle - self.warm) / max(self.cur - self.warm, 1)
                    lr = self.min_lr + (self.max_lr - self.min_lr) * (1 + math.cos(math.pi * t)) / 2
                return [lr for _ in self.optimizer.param_groups]

            def step(self, epoch=None):
                if epoch is None:
                    epoch = self.last_epoch + 1
                    self.step_in_cycle += 1
                    if self.step_in_cycle >= self.cur:
                        self.cycle += 1
                        self.step_in_cycle -= self.cur
                        self.cur = int((self.cur - self.warm) * self.mult) + self.warm
                else:
                    if epoch >= self.first:
                        if self.mult == 1.0:
                            self.step_in_cycle, self.cycle = epoch % self.first, epoch // self.first
                        else:
                            n = int(math.log(epoch / self.first * (self.mult - 1) + 1, self.mult))
                            self.cycle = n
                            self.step_in_cycle = epoch - int(self.first * (self.mult ** n - 1) / (self.mult - 1))
                            self.cur = self.first * self.mult ** n
                    else:
                        self.cur, self.step_in_cycle = self.first, epoch
                self.max_lr = self.base_max * (self.gamma ** self.cycle)
                self.last_epoch = math.floor(epoch)
                for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
                    g["lr"] = lr

        caw = types.ModuleType("cosine_annealing_warmup")
        caw.CosineAnnealingWarmupRestarts = CosineAnnealingWarmupRestarts
        sys.modules["cosine_annealing_warmup"] = caw


_imported = {}


def import_reference():
    """Returns a namespace with the reference's hot-path classes (imported once)."""
    if _imported:
        return types.SimpleNamespace(**_imported)
    root = find_reference()
    if root is None:
        raise RuntimeError("reference sources not found (looked in %s)" % REFERENCE_CANDIDATES)
    _install_stubs()
    sys.path.insert(0, root)
    import transformers

    transformers.BertModel.from_pretrained = classmethod(
        lambda cls, *a, **k: transformers.BertModel(transformers.BertConfig()))
    from lrce.feature_extractor import video_swin_ori as swin
    from lrce.feature_extractor.video import VideoExtractor
    from lrce.models import e2e, embedding, fusionv3

    _imported.update(root=root, swin=swin, VideoExtractor=VideoExtractor, e2e=e2e, embedding=embedding,
                     fusionv3=fusionv3)
    return types.SimpleNamespace(**_imported)


def import_agents():
    """the reference's agents (lrce/agent/*.py: AgentOE / AgentMC / AgentCount), imported unmodified"""
    import_reference()
    from lrce.agent import AgentCount, AgentMC, AgentOE

    return types.SimpleNamespace(AgentOE=AgentOE, AgentMC=AgentMC, AgentCount=AgentCount)


@contextlib.contextmanager
def reference_workdir(swin_state_dict=None):
    """chdir into a scratch dir that holds ./pretrained_models/<swin ckpt> in the layout video.py:20-26 expects
    ({'state_dict': {'backbone.' + key: tensor}}), so that E2E* constructors run unmodified (e2e.py:11-14)."""
    import torch

    ref = import_reference()
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "pretrained_models"))
        if swin_state_dict is None:
            torch.manual_seed(0)
            m = ref.swin.SwinTransformer3D(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32],
                                           patch_size=(2, 4, 4), window_size=(8, 7, 7), drop_path_rate=0.2,
                                           patch_norm=True)
            swin_state_dict = m.state_dict()
        torch.save({"state_dict": {"backbone." + k: v for k, v in swin_state_dict.items()}},
                   os.path.join(tmp, "pretrained_models", "swin_base_patch244_window877_kinetics600_22k.pth"))
        os.chdir(tmp)
        try:
            yield tmp
        finally:
            os.chdir(old)


def build_reference_e2e(kind, cfg, state_dict=None):
    """kind in {'oe','mc','count'}; cfg = dict(feature_dim, num_classes, video_feature_res, video_feature_dim,
    frame_sample_size, temporal_scale, text_seq_len). Optionally loads `state_dict` (strict)."""
    ref = import_reference()
    cls = {"oe": ref.e2e.E2EOpenEnded, "mc": ref.e2e.E2EMultipleChoice, "count": ref.e2e.E2ECount}[kind]
    swin_sd = None
    if state_dict is not None:
        pre = "video_extractor.swin."
        swin_sd = {k[len(pre):]: v for k, v in state_dict.items() if k.startswith(pre)}
    with reference_workdir(swin_sd):
        model = cls(**cfg)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    return model
