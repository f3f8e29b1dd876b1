"""tf.app.flags look-alike: `--flag value` pairs, booleans given as the separate
tokens True/False, unknown flags silently ignored (the shipped launchers rely on that:
run_hmf.sh passes --vocab_min_thresh, run_lstm.sh passes --steps_per_checkpoint)."""
import argparse


def _bool(v):
    if isinstance(v, bool):
        return v
    return str(v).lower() in ('true', '1', 't', 'yes', 'y')


class Flags(object):
    def __init__(self):
        self._parser = argparse.ArgumentParser(allow_abbrev=False)
        self._parsed = None

    def DEFINE_string(self, name, default, doc=''):
        self._parser.add_argument('--' + name, type=str, default=default, help=doc)

    def DEFINE_integer(self, name, default, doc=''):
        self._parser.add_argument('--' + name, type=int, default=default, help=doc)

    def DEFINE_float(self, name, default, doc=''):
        self._parser.add_argument('--' + name, type=float, default=default, help=doc)

    def DEFINE_boolean(self, name, default, doc=''):
        self._parser.add_argument('--' + name, type=_bool, default=default, nargs='?', const=True, help=doc)

    def parse(self, argv=None):
        self._parsed, unknown = self._parser.parse_known_args(argv)
        return unknown

    def __getattr__(self, k):
        if k.startswith('_'):
            raise AttributeError(k)
        if self._parsed is None:
            self.parse()
        return getattr(self._parsed, k)

    def __setattr__(self, k, v):
        if k.startswith('_'):
            object.__setattr__(self, k, v)
        else:
            if self._parsed is None:
                self.parse()
            setattr(self._parsed, k, v)
