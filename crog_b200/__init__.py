"""crog_b200 — B200-native (sm_100a) implementation of CROG's batched referring-grasp inference path."""
__all__ = ["CROG", "build_crog"]


def __getattr__(name):  # lazy: importing the package must not require torch/CUDA
    if name in __all__:
        from . import model

        return getattr(model, name)
    raise AttributeError(name)
