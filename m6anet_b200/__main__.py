"""Command line: `python -m m6anet_b200 inference ...` (the `m6anet inference` sub-command of the reference,
m6anet/__init__.py:11-30).  Sub-command = module exposing argparser() + main(args), as in the reference."""
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser

from . import __version__, inference

modules = ['inference']
_NOT_PORTED = ['dataprep', 'train', 'compute_norm_factors', 'convert']


def main(argv=None):
    parser = ArgumentParser(prog='m6anet', formatter_class=ArgumentDefaultsHelpFormatter)
    parser.add_argument('-v', '--version', action='version', version='%(prog)s-b200 {version}'.format(version=__version__))
    subparsers = parser.add_subparsers(title='subcommands', description='valid commands', help='additional help',
                                       dest='command')
    subparsers.required = True
    for module in modules:
        mod = globals()[module]
        p = subparsers.add_parser(module, parents=[mod.argparser()])
        p.set_defaults(func=mod.main)
    for name in _NOT_PORTED:
        p = subparsers.add_parser(name, help='not part of m6anet_b200 (inference hot path only); use the reference m6anet')
        p.set_defaults(func=lambda a, n=name: parser.error(
            f"'{n}' is outside the m6anet_b200 scope (B200 inference hot path only); run it with the reference m6anet"))
    args = parser.parse_args(argv)
    args.func(args)


if __name__ == "__main__":
    main()
