"""Condenses a compute-sanitizer racecheck log into unique (access A, access B) source-line pairs with the source text
of both lines (run in the same tree the log was produced from).  usage: race_triage.py log [log ...]"""
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pat = re.compile(r"(Write|Read) access at .*? in (\S+?):(\d+)(?: \[(\d+) hazards\])?")
src_cache = {}


def line_text(fn, ln):
    path = os.path.join(ROOT, "locator_b200", "csrc", os.path.basename(fn))
    if path not in src_cache:
        try:
            src_cache[path] = open(path).read().splitlines()
        except OSError:
            src_cache[path] = []
    lines = src_cache[path]
    return lines[ln - 1].strip()[:110] if 0 < ln <= len(lines) else "?"


for log in sys.argv[1:]:
    pairs = collections.OrderedDict()
    cur = None
    summary = ""
    for raw in open(log, errors="replace"):
        if "SUMMARY" in raw:
            summary = raw.strip("= \n")
        m = pat.search(raw)
        if not m:
            continue
        kind, fn, ln, hz = m.group(1), m.group(2), int(m.group(3)), m.group(4)
        if "Race reported" in raw:
            cur = (kind, fn, ln)
        elif cur is not None:
            key = (cur, (kind, fn, ln))
            pairs[key] = pairs.get(key, 0) + int(hz or 0)
            cur = None
    print(f"## {os.path.basename(log)}: {summary}; {len(pairs)} distinct pairs\n")
    for ((ka, fa, la), (kb, fb, lb)), hz in pairs.items():
        print(f"* {ka} `{os.path.basename(fa)}:{la}` `{line_text(fa, la)}`  \n  vs {kb} `{os.path.basename(fb)}:{lb}` `{line_text(fb, lb)}`  ({hz} hazards)")
    print()
