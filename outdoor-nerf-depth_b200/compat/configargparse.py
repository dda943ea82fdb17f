"""Stand-in for the subset of ``configargparse`` the reference trainer uses (ddp_train_nerf.py:657-727):
``ArgumentParser`` whose ``--config`` option (``is_config_file=True``) names a ``key = value`` file supplying defaults that
the command line overrides, plus ``format_values()``.  Used only when the real package is not installed."""
import argparse


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._config_dests = []
        self._source = {}

    def add_argument(self, *names, **kw):
        is_cfg = kw.pop("is_config_file", False)
        act = super().add_argument(*names, **kw)
        if is_cfg:
            self._config_dests.append(act.dest)
        return act

    @staticmethod
    def _read(path):
        items = []
        with open(path) as f:
            for line in f:
                line = line.split("#", 1)[0].strip()
                if not line or line.startswith("["):
                    continue
                key, sep, val = line.partition("=")
                if not sep:
                    key, _, val = line.partition(" ")
                items.append((key.strip().lstrip("-"), val.strip().strip("'\"")))
        return items

    def parse_known_args(self, args=None, namespace=None):
        import sys
        args = list(sys.argv[1:] if args is None else args)
        # first pass: only to find the config file option(s)
        pre, _ = super().parse_known_args(args, None)
        file_args = []
        by_dest = {a.dest: a for a in self._actions}
        for dest in self._config_dests:
            path = getattr(pre, dest, None)
            if not path or str(path) == "None":
                continue
            for key, val in self._read(path):
                act = by_dest.get(key)
                if act is None or key in self._config_dests:
                    continue
                opt = act.option_strings[0]
                if isinstance(act, (argparse._StoreTrueAction, argparse._StoreFalseAction)):
                    if val.lower() in ("true", "1", "yes") and isinstance(act, argparse._StoreTrueAction):
                        file_args.append(opt)
                    elif val.lower() in ("false", "0", "no") and isinstance(act, argparse._StoreFalseAction):
                        file_args.append(opt)
                elif val != "None":
                    file_args += [opt, val]
                self._source[key] = "config file"
        ns, rest = super().parse_known_args(file_args + args, namespace)       # later (command-line) values win
        return ns, rest

    def format_values(self):
        return "configargparse stand-in: values from %s" % (sorted(set(self._source.values())) or ["defaults / command line"])


ArgParser = ArgumentParser
