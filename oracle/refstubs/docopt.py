"""TEST INFRASTRUCTURE ONLY -- import-time stand-in for the absent `docopt` wheel (the reference's train.py scripts import it at
module level; the tests import those modules for their model factories and never parse a command line)."""


def docopt(*args, **kwargs):
    raise RuntimeError("docopt stub: command-line parsing is not available in the test environment")
