"""Shared helpers for the test-suite: package / oracle loading and problem fixtures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def package():
    libs = [os.path.join(entry.PKG_DIR, "libgomp_host.so"), os.path.join(entry.PKG_DIR, "libgomp_cuda.so")]
    if not all(os.path.exists(p) for p in libs):
        entry.build()
    return entry.load_package()


def oracle():
    return entry.load_oracle()
