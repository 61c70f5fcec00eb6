"""Mirror of the reference's ``model`` package factory (model/__init__.py:6-28)."""
from .crog import CROG
from .ssg import SSG


def build_crog(args):
    """Returns ``(model, param_list)`` like the reference; the two lr groups (backbone vs head) are
    kept for signature compatibility although this implementation is inference only."""
    model = CROG(args)
    backbone, head = [], []
    for k, v in model.named_parameters():
        (backbone if (k.startswith("backbone") and "positional_embedding" not in k) else head).append(v)
    lr_multi, base_lr = getattr(args, "lr_multi", 0.1), getattr(args, "base_lr", 1e-4)
    return model, [{"params": backbone, "initial_lr": lr_multi * base_lr}, {"params": head, "initial_lr": base_lr}]


def build_ssg(args):
    """model/__init__.py:26-29: returns ``(model, model.parameters())``."""
    model = SSG(args)
    return model, model.parameters()
