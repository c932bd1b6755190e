"""Minimal run log (reference infolog.py): timestamped lines to stdout and to <log_dir>/Terminal_train_log."""
import atexit
from datetime import datetime

_file = None


def init(filename, run_name, slack_url=None):
    global _file
    _close()
    _file = open(filename, "a", encoding="utf-8")
    _file.write("\n-----------------------------------------------------------------\nStarting new %s training run\n" % run_name)


def log(msg, end="\n", slack=False):
    print(msg, end=end)
    if _file is not None:
        _file.write("[%s]  %s\n" % (datetime.now().strftime("%Y-%m-%d %H:%M:%S.%f")[:-3], msg))
        _file.flush()


def _close():
    global _file
    if _file is not None:
        _file.close()
        _file = None


atexit.register(_close)
