"""Host-side post-processing of the generated text, mirroring ``ask()`` (Vidi1.5_9B/vidi/eval/inference.py:52-66,
Vidi_7B/inference.py:51-65): normalised ``start-end`` fractions -> ``HH:MM:SS-HH:MM:SS`` strings."""
from __future__ import annotations

import re
from typing import List, Tuple

_PATTERN = re.compile(r"(\d\.\d+)-(\d\.\d+)")                 # inference.py:55


def parse_ranges(text: str) -> List[Tuple[float, float]]:
    return [(float(a), float(b)) for a, b in _PATTERN.findall(text.strip())]


def format_time_ranges(text: str, length_s: float) -> str:
    """Exactly the arithmetic of inference.py:59-65 (int() truncation, hours from t/3600, minutes from (int(t)%3600)//60)."""
    out = []
    for a, b in parse_ranges(text):
        t0, t1 = a * length_s, b * length_s
        out.append("{:02d}:{:02d}:{:02d}-{:02d}:{:02d}:{:02d}".format(
            int(t0 / 3600), (int(t0) % 3600) // 60, int(t0) % 60, int(t1 / 3600), (int(t1) % 3600) // 60, int(t1) % 60))
    return ", ".join(out)
