"""Stand-in for ``tensorboardX.SummaryWriter`` (ddp_train_nerf.py:13,413,511,624-638): scalars are appended to
``<logdir>/scalars.jsonl`` instead of an event file.  Used only when the real package is not installed."""
import json
import os


class SummaryWriter(object):
    def __init__(self, logdir=None, *a, **k):
        self.logdir = logdir
        self._f = None
        if logdir:
            os.makedirs(logdir, exist_ok=True)
            self._f = open(os.path.join(logdir, "scalars.jsonl"), "a")

    def add_scalar(self, tag, value, global_step=None, *a, **k):
        if self._f:
            self._f.write(json.dumps({"tag": tag, "value": float(value), "step": global_step}) + "\n")
            self._f.flush()

    def add_image(self, *a, **k):
        pass

    def close(self):
        if self._f:
            self._f.close()
            self._f = None
