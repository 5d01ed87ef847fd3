"""Timestamped record log carried by the analyzer objects (host-side bookkeeping only).

Same surface as the reference's ``InfoManagementSystem``
(/root/reference/src/fiberis/utils/history_utils.py:11-181): ``records`` is a list of
``{'description', 'timestamp', 'level'}`` dicts.
"""
from __future__ import annotations

import datetime
import os
from typing import Any, Callable, Dict, List, Optional

_STAMP = "%Y-%m-%d %H:%M:%S.%f"


def _fmt(ts: datetime.datetime) -> str:
    return ts.strftime(_STAMP)[:-3]


class InfoManagementSystem:
    def __init__(self):
        self.records: List[Dict[str, Any]] = []

    def add_record(self, description, level="INFO") -> None:
        self.records.append({
            "description": description if isinstance(description, str) else str(description),
            "timestamp": datetime.datetime.now(),
            "level": str(level).upper(),
        })

    def get_records(self, filter_fn: Optional[Callable[[Dict[str, Any]], bool]] = None) -> List[Dict[str, Any]]:
        return [r for r in self.records if filter_fn is None or filter_fn(r)]

    def print_records(self, filter_fn=None) -> None:
        recs = self.get_records(filter_fn)
        if not recs:
            print("No records to display.")
        for r in recs:
            print(f"[{r.get('level', 'N/A')}] {_fmt(r['timestamp'])}: {r['description']}")

    def save_records_to_txt(self, save_path: Optional[str] = None, filter_fn=None) -> str:
        recs = self.get_records(filter_fn)
        if not recs:
            print("No records to save.")
            return ""
        auto = f"history_log_{datetime.datetime.now().strftime('%Y-%m-%d_%H-%M-%S')}.txt"
        if save_path and os.path.isdir(save_path):
            path = os.path.join(save_path, auto)
        elif save_path:
            path = save_path
        else:
            path = os.path.join(os.getcwd(), auto)
        try:
            with open(path, "w") as fh:
                for r in recs:
                    fh.write(f"Timestamp: {_fmt(r['timestamp'])}\nLevel:     {r.get('level', 'N/A')}\n"
                             f"Message:   {r['description']}\n{'-' * 50}\n")
        except IOError as exc:
            print(f"Error saving records to {path}: {exc}")
            return ""
        print(f"Records successfully saved to {path}")
        return path

    def clear_records(self) -> None:
        self.records = []

    def __repr__(self) -> str:
        return f"InfoManagementSystem(records_count={len(self.records)})"

    def __str__(self) -> str:
        if not self.records:
            return "No records in history."
        lines = [f"History Log ({len(self.records)} records):"]
        lines += [f"  [{r.get('level', 'N/A')}] {_fmt(r['timestamp'])}: {r['description']}" for r in self.records]
        return "\n".join(lines)
