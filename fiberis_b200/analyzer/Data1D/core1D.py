"""``Data1D`` — the source-term interface of the simulator (boundary subset of the reference class).

The simulator reads ``taxis``, ``start_time`` and ``get_value_by_time`` only
(/root/reference/src/fiberis/analyzer/Data1D/core1D.py:15-48, 202-248); npz load/save are kept so
gauge files written by the reference can be used as sources.  Signal-processing methods
(crop/filter/merge/plot ...) are outside the hot path and not provided.
"""
from __future__ import annotations

import datetime
import os
from typing import Optional, Union

import numpy as np

from fiberis_b200.utils import history_utils


class Data1D:
    def __init__(self, data: Optional[np.ndarray] = None, taxis: Optional[np.ndarray] = None,
                 start_time: Optional[datetime.datetime] = None, name: Optional[str] = None):
        self.taxis = taxis
        self.data = data
        self.start_time = start_time
        self.name = name
        self.history = history_utils.InfoManagementSystem()
        self.history.add_record(f"Initialized Data1D object with name: {name}" if name
                                else "Initialized empty Data1D object.")

    # -- setters -------------------------------------------------------------------------------
    def set_data(self, data: np.ndarray) -> None:
        self.data = data
        self.history.add_record("Data attribute updated.")

    def set_taxis(self, taxis: np.ndarray) -> None:
        self.taxis = taxis
        self.history.add_record("Time axis (taxis) updated.")

    def set_start_time(self, start_time: datetime.datetime) -> None:
        self.start_time = start_time
        self.history.add_record(f"Start time set to {start_time}.")

    def set_name(self, name: str) -> None:
        self.name = name
        self.history.add_record(f"Name set to '{name}'.")

    # -- I/O -----------------------------------------------------------------------------------
    def load_npz(self, filename: str) -> None:
        """Keys 'data', 'taxis', 'start_time' (datetime or ISO string); values cast to float."""
        if not os.path.exists(filename):
            self.history.add_record(f"Error: File not found at {filename}", level="ERROR")
            raise FileNotFoundError(f"The specified NPZ file does not exist: {filename}")
        with np.load(filename, allow_pickle=True) as z:
            missing = [k for k in ("data", "taxis", "start_time") if k not in z]
            if missing:
                raise KeyError(f"Missing keys in NPZ file: {', '.join(missing)}")
            data = z["data"].astype(float)
            taxis = z["taxis"].astype(float)
            raw = z["start_time"].item()
        if isinstance(raw, str):
            start = datetime.datetime.fromisoformat(raw)
        elif isinstance(raw, datetime.datetime):
            start = raw
        else:
            raise ValueError(f"Unsupported start_time type in NPZ file: {type(raw)}")
        self.data, self.taxis, self.start_time = data, taxis, start
        self.name = os.path.basename(filename)
        self.history.add_record(f"Successfully loaded data from {filename}. Name set to '{self.name}'.")

    def savez(self, filename: str) -> None:
        if self.data is None or self.taxis is None:
            raise ValueError("Data and taxis must be set before saving.")
        if not filename.endswith(".npz"):
            filename += ".npz"
        np.savez(filename, data=self.data, taxis=self.taxis, start_time=self.start_time)
        self.history.add_record(f"Data successfully saved to {filename}.")

    # -- the one method on the simulator's path ---------------------------------------------------
    def get_value_by_time(self, time_point: Union[datetime.datetime, float, int]) -> float:
        """Linear interpolation of ``data`` at ``time_point`` (seconds from ``start_time`` or an
        absolute datetime), clamped to the end values outside ``taxis`` (``np.interp``)."""
        if self.data is None or self.taxis is None or self.data.size == 0 or self.taxis.size == 0:
            self.history.add_record(f"Error: Cannot get value at {time_point}, data/taxis is not loaded or empty.",
                                    level="ERROR")
            raise ValueError("Data or taxis is not loaded or is empty.")
        if isinstance(time_point, (int, float)):
            seconds = float(time_point)
        elif isinstance(time_point, datetime.datetime):
            if self.start_time is None:
                raise ValueError("start_time is not set when using a datetime object for time_point.")
            seconds = (time_point - self.start_time).total_seconds()
        else:
            raise TypeError("time_point must be either datetime.datetime or float/int (seconds).")
        return float(np.interp(seconds, self.taxis, self.data))

    def __str__(self) -> str:
        n = 0 if self.data is None else len(self.data)
        return f"Data1D(name={self.name!r}, points={n}, start_time={self.start_time})"
