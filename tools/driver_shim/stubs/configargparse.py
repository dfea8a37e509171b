"""Minimal configargparse stand-in (absent from this image): argparse plus one `is_config_file` option whose file holds
`key = value` lines and bare flag names (the format of the reference's configs/*_lushnerf).  Command-line arguments
override the file, as in configargparse."""
import argparse
import sys


class ArgumentParser(argparse.ArgumentParser):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._config_opts = []

    def add_argument(self, *names, **kw):
        if kw.pop("is_config_file", False):
            self._config_opts += [n for n in names if n.startswith("-")]
        return super().add_argument(*names, **kw)

    def _file_args(self, path):
        out = []
        for line in open(path):
            line = line.split("#", 1)[0].strip()
            if not line:
                continue
            if "=" in line:
                k, v = (t.strip() for t in line.split("=", 1))
                act = self._option_string_actions.get("--" + k)
                if act is not None and act.nargs == 0:        # store_true / store_false written as `flag = True`
                    if v.lower() in ("true", "1", "yes"):
                        out.append("--" + k)
                else:
                    out += ["--" + k, v]
            else:
                out.append("--" + line)
        return out

    def parse_known_args(self, args=None, namespace=None):
        args = list(sys.argv[1:] if args is None else args)
        extra = []
        for i, a in enumerate(args):
            for opt in self._config_opts:
                if a == opt and i + 1 < len(args):
                    extra += self._file_args(args[i + 1])
                elif a.startswith(opt + "="):
                    extra += self._file_args(a.split("=", 1)[1])
        return super().parse_known_args(extra + args, namespace)   # later (command-line) values win
