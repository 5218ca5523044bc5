"""Command line of the reference (vlgp/__main__.py:6-22): python -m vlgp_b200 <input> <output> <n_factors>."""
import argparse

from . import api, util


def cli(argv=None):
    ap = argparse.ArgumentParser(prog="vlgp_b200", description="variational Latent Gaussian Process (vLGP) on B200")
    ap.add_argument("fin", metavar="<path to input file>")
    ap.add_argument("fout", metavar="<path to output file>")
    ap.add_argument("n_factors", type=int, metavar="<number of factors>")
    ap.add_argument("--max_iter", type=int, default=20, help="Maximum number of iterations")
    ap.add_argument("--min_iter", type=int, default=5, help="Minimum number of iterations")
    args = ap.parse_args(argv)
    print("Loading {}".format(args.fin))
    trials = util.load(args.fin)
    trials = list(trials) if not isinstance(trials, dict) else trials.get("trials", trials)
    result = api.fit(trials, args.n_factors, max_iter=args.max_iter, min_iter=args.min_iter, path=args.fout)
    print("Saving {}".format(args.fout))
    util.save(result, args.fout)


if __name__ == "__main__":
    cli()
