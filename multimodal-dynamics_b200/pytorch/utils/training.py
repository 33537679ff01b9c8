"""Host-side helpers used around the step (mirror of mmdyn/pytorch/utils/training.py): a progress
line that works without a TTY (the reference runs `stty size` at import and fails headless), and
pickle helpers."""
import pickle
import sys
import time

_t0 = time.time()
_last = _t0


def format_time(seconds):
    seconds = int(seconds)
    h, r = divmod(seconds, 3600)
    m, s = divmod(r, 60)
    return (f"{h}h" if h else "") + (f"{m}m" if m or h else "") + f"{s}s"


def progress_bar(current, total, msg=None):
    global _t0, _last
    now = time.time()
    if current <= 1:
        _t0 = now
    step, tot = now - _last, now - _t0
    _last = now
    line = f"\r[{current}/{total}] step {format_time(step)} | tot {format_time(tot)}"
    if msg:
        line += " | " + str(msg)
    sys.stdout.write(line + ("\n" if current >= total else ""))
    sys.stdout.flush()


def save_pkl(obj, path):
    with open(path, "wb") as f:
        pickle.dump(obj, f, pickle.HIGHEST_PROTOCOL)


def load_pkl(path):
    with open(path, "rb") as f:
        return pickle.load(f)
