"""Minimal stand-in for the reference's aparse.ArgParser (aparse/__init__.py:13-46): argparse groups chained
through get_parser(ps), parsed once into a dict. Only what Synthesizer.get_parser needs."""
import argparse


class ArgParser(object):
    def __init__(self, ps=None, name=None, **kwargs):
        if ps is None:
            self._ps = argparse.ArgumentParser(**kwargs)
        elif isinstance(ps, ArgParser):
            self._ps = ps._ps
        else:
            self._ps = ps
        self._group = self._ps.add_argument_group(name) if name else self._ps

    def add(self, *args, **kwargs):
        self._group.add_argument(*args, **kwargs)
        return self

    def add_flag(self, name, **kwargs):
        self._group.add_argument(name, action="store_true", **kwargs)
        return self

    def parse_args(self, args=None):
        return vars(self._ps.parse_args(args))
