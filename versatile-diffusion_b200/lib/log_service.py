"""print_log stand-in (reference lib/log_service.py:15-36: rank-0 console print)."""
import os


def print_log(*console_info):
    if int(os.environ.get("RANK", "0")) == 0:
        print(" ".join(str(i) for i in console_info))
