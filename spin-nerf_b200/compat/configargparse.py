"""Minimal configargparse stand-in (the image has no configargparse; DS_NeRF/run_nerf.py:741-925 needs it):
argparse + `--config file` with `key = value` lines; CLI flags override the file."""
import argparse


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *a, **k):
        k.pop("default_config_files", None)
        super().__init__(*a, **k)
        self._cfg_dest = None

    def add_argument(self, *a, **k):
        if k.pop("is_config_file", False):
            k.setdefault("default", None)
            act = super().add_argument(*a, **k)
            self._cfg_dest = act.dest
            return act
        return super().add_argument(*a, **k)

    def parse_known_args(self, args=None, namespace=None):
        import sys
        args = list(sys.argv[1:] if args is None else args)
        pre, _ = super().parse_known_args(args, None)
        cfg = getattr(pre, self._cfg_dest, None) if self._cfg_dest else None
        if cfg:
            file_args = []
            opts = {s for act in self._actions for s in act.option_strings}
            for line in open(cfg):
                line = line.split("#", 1)[0].strip()
                if not line:
                    continue
                key, _, val = line.partition("=")
                key, val = "--" + key.strip(), val.strip()
                if key not in opts:
                    continue
                act = next(a_ for a_ in self._actions if key in a_.option_strings)
                if isinstance(act, (argparse._StoreTrueAction, argparse._StoreFalseAction)):
                    if val.lower() in ("true", "1", "yes", ""):
                        file_args.append(key)
                else:
                    if len(val) >= 2 and val[0] == val[-1] and val[0] in "\"'":
                        val = val[1:-1]
                    file_args += [key] + (val.strip("[]").replace(",", " ").split() if act.nargs in ("+", "*") else [val])
            args = file_args + args
        return super().parse_known_args(args, namespace)


ArgParser = ArgumentParser
