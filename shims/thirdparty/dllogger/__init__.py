"""Minimal stand-in for NVIDIA/dllogger (absent offline; unpinned git dependency of the reference, Dockerfile:24): the subset of
the API run_pretraining.py / run_squad.py call -- init, log, flush, metadata, Verbosity, JSONStreamBackend, StdOutBackend.
JSON lines carry the same "DLLL " prefix and {type, step, data} layout so existing log scrapers keep working."""
import atexit
import json
import os
import time


class Verbosity:
    OFF, DEFAULT, VERBOSE = -1, 0, 1


class JSONStreamBackend:
    def __init__(self, verbosity=Verbosity.DEFAULT, filename="dllogger.json", append=False):
        d = os.path.dirname(os.path.abspath(filename))
        os.makedirs(d, exist_ok=True)
        self.f = open(filename, "a" if append else "w")
        atexit.register(self.f.close)

    def log(self, rec):
        self.f.write("DLLL " + json.dumps(rec, default=str) + "\n")

    def flush(self):
        self.f.flush()


class StdOutBackend:
    def __init__(self, verbosity=Verbosity.DEFAULT, step_format=None, metric_format=None, prefix_format=None):
        self.step_format = step_format or (lambda s: str(s))

    def log(self, rec):
        if rec["type"] != "LOG":
            return
        print("DLL {} - {} {}".format(rec["datetime"], self.step_format(rec["step"]),
                                      " ".join("{} : {}".format(k, v) for k, v in rec["data"].items())), flush=True)

    def flush(self):
        pass


_backends = []


def init(backends):
    global _backends
    _backends = list(backends)


def _emit(rec):
    for b in _backends:
        b.log(rec)


def log(step, data, verbosity=Verbosity.DEFAULT):
    _emit({"type": "LOG", "datetime": time.strftime("%Y-%m-%d %H:%M:%S"), "elapsedtime": time.perf_counter(),
           "step": list(step) if isinstance(step, tuple) else step, "data": data})


def metadata(metric, meta):
    _emit({"type": "METADATA", "metric": metric, "metadata": meta})


def flush():
    for b in _backends:
        b.flush()
