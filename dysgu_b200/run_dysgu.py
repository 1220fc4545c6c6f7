"""Run dysgu's command line with the alignment path on the GPU engine:

    python -m dysgu_b200.run_dysgu call ref.fa wd input.bam ...      # = dysgu call ..., alignments batched on the GPU
    python -m dysgu_b200.run_dysgu --stock call ...                  # the unmodified CPU path (for side-by-side timing)

Before handing over to `dysgu.main:cli` (pyproject.toml:40-41) it routes dysgu's two alignment modules to this package
(dysgu_b200.install) and wraps the alignment loops with caller-side batching (dysgu_b200.patch.install).  With
DB200_CAPTURE=<file> every alignment and its result is recorded (dysgu_b200.capture) - in --stock mode that is the
oracle file scripts/replay_capture.py replays.  Needs an importable dysgu."""
import os
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    stock = "--stock" in argv
    if stock:
        argv.remove("--stock")
    if not stock:
        from . import patch
        done = patch.install()
        if os.environ.get("DB200_DEBUG"):
            print("[db200] patched:", ", ".join(done), file=sys.stderr)
    session = None
    if os.environ.get("DB200_CAPTURE"):
        from . import capture
        session = capture.install(os.environ["DB200_CAPTURE"])
    from dysgu.main import cli
    sys.argv = ["dysgu"] + argv
    try:
        cli()
    finally:
        if session is not None:
            session.close()


if __name__ == "__main__":
    main()
