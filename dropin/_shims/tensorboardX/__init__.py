"""Minimal stand-in for tensorboardX (used only when the real package is not installed)."""
try:
    from torch.utils.tensorboard import SummaryWriter  # noqa: F401  (same interface)
except Exception:  # noqa: BLE001  (tensorboard itself missing)
    class SummaryWriter:  # type: ignore[no-redef]
        def __init__(self, *args, **kwargs):
            pass

        def add_scalar(self, *args, **kwargs):
            pass

        def close(self):
            pass
