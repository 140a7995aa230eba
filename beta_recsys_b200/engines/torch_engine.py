"""Host-side mirror of the reference's ``ModelEngine`` (beta_rec/models/torch_engine.py).

Same constructor contract (``Engine(config)``), same attributes and methods the
callers use -- ``set_device``, ``set_optimizer``, ``save_checkpoint``,
``resume_checkpoint``, ``bpr_loss``, ``bce_loss``, ``writer`` -- but the
optimizer is a descriptor for the CUDA row-update kernels instead of a
``torch.optim`` object, and the device MUST be CUDA: there is no CPU path.
"""
import torch
import torch.nn.functional as F

from .. import _lib


class _NullWriter(object):
    """Stands in for tensorboardX.SummaryWriter when tensorboardX is absent."""

    def add_scalar(self, *a, **k):
        pass

    def add_scalars(self, *a, **k):
        pass

    def add_text(self, *a, **k):
        pass

    def close(self):
        pass


def make_writer(log_dir):
    try:  # the reference hard-requires tensorboardX (torch_engine.py:3); keep it optional here
        from tensorboardX import SummaryWriter

        return SummaryWriter(log_dir=log_dir)
    except Exception:
        return _NullWriter()


class RowOptimizer(object):
    """What ``engine.optimizer`` is in this build: the optimizer kind, its
    hyper-parameters (torch.optim defaults) and, for Adam/RMSprop, the state
    tensors keyed by parameter name.  The step itself runs in libbrs_b200
    (csrc/rows_apply.cu).

    mode ``"dense"`` (default) reproduces the reference: dense gradients make
    torch.optim.Adam/RMSprop update EVERY row each step.  mode ``"touched"``
    updates only the rows present in the batch (lazy Adam) -- NOT equivalent to
    the reference after the first step; opt-in via config["model"]["adam_mode"].
    """

    def __init__(self, kind, lr, mode="dense"):
        if kind not in _lib.OPT_KINDS:
            raise ValueError("unsupported optimizer %r (sgd | adam | rmsprop)" % (kind,))
        if mode not in ("dense", "touched"):
            raise ValueError("adam_mode must be 'dense' or 'touched'")
        self.kind = kind
        self.lr = float(lr)
        self.mode = mode
        self.state = {}  # name -> {"m": tensor, "v": tensor}
        self.desc = _lib.make_opt(kind, lr, _lib.DENSE if mode == "dense" else _lib.TOUCHED_ROWS)

    def add_param(self, name, tensor):
        st = {}
        if self.kind == "adam":
            st["m"] = torch.zeros_like(tensor)
        if self.kind in ("adam", "rmsprop"):
            st["v"] = torch.zeros_like(tensor)
        self.state[name] = st
        return st

    def zero_grad(self):  # gradients live in kernel-owned scratch that is cleared by the step itself
        pass


class ModelEngine(object):
    """Mirror of beta_rec.models.torch_engine.ModelEngine.  Subclasses set ``self.model``."""

    def __init__(self, config):
        self.config = config
        self.set_device()
        self.set_optimizer()
        self.model.to(self.device)
        self.writer = make_writer(config["system"]["run_dir"] if "system" in config else None)

    def set_optimizer(self):
        """torch_engine.py:23-39 -- sgd / adam / rmsprop with lr from config["model"]["lr"]."""
        m = self.config["model"]
        if m["optimizer"] not in ("sgd", "adam", "rmsprop"):
            raise ValueError("unsupported optimizer %r" % (m["optimizer"],))
        mode = m["adam_mode"] if "adam_mode" in m else "dense"
        self.optimizer = RowOptimizer(m["optimizer"], m["lr"], mode)

    def set_device(self):
        """torch_engine.py:41-45.  A non-CUDA device is an error here, loudly."""
        self.device = torch.device(self.config["model"]["device_str"])
        if self.device.type != "cuda":
            raise _lib.BrsError(
                "beta_recsys_b200 engines run on a CUDA (sm_100a) device only; got device_str=%r. "
                "There is no CPU fallback -- use the reference engine for CPU runs."
                % (self.config["model"]["device_str"],)
            )
        if not torch.cuda.is_available():
            raise _lib.BrsError("CUDA device requested but torch.cuda.is_available() is False")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _lib.load()  # fail now, not at the first batch, if the library is missing
        self.model.device = self.device

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _check_predict(self):
        """The no_grad scoring kernels raise their OWN error word (brs_step_ws.predict_err, byte 44 of the
        workspace) for an out-of-range id -- never the training step's -- and publish NaN for that sample.
        Reading it costs one 4-byte D2H; predict's callers move the scores to the host right after
        (core/eval_engine.py:258-273).  The reference raises IndexError inside nn.Embedding."""
        word = self._ws[44:48].view(torch.int32)
        if int(word.item()):
            word.zero_()
            raise IndexError("index out of range in self")

    def save_checkpoint(self, model_dir):
        """torch_engine.py:70-73 -- same state_dict keys/shapes as the reference module."""
        assert hasattr(self, "model"), "Please specify the exact model !"
        torch.save(self.model.state_dict(), model_dir)

    def resume_checkpoint(self, model_dir, model=None):
        """torch_engine.py:76-90."""
        assert hasattr(self, "model"), "Please specify the exact model !"
        print("loading model from:", model_dir)
        state_dict = torch.load(model_dir, map_location=self.device)
        target = self.model if model is None else model
        # copy INTO the existing storage: the kernels hold raw pointers to it
        with torch.no_grad():
            own = target.state_dict()
            missing = set(own) ^ set(state_dict)
            if missing:
                raise RuntimeError("checkpoint keys do not match the model: %s" % sorted(missing))
            for k, v in state_dict.items():
                own[k].copy_(v)
        return target

    def bpr_loss(self, pos_scores, neg_scores):
        """torch_engine.py:92-106 (API parity; training uses the fused kernel)."""
        return -torch.mean(F.logsigmoid(pos_scores - neg_scores))

    def bce_loss(self, scores, ratings):
        """torch_engine.py:108-121 (API parity; training uses the fused kernel)."""
        return torch.nn.BCELoss()(scores, ratings)
