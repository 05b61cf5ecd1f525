"""Portable run-parameter / runtime-metrics sink (second half).

Replaces ``Assets/Scripts/SQL.cs`` of the reference, which inserts into a SQLite file at a hard-coded Windows path
(``SQL.cs:58``, ``:105``) through Mono.Data.Sqlite + ``Assets/Plugin/sqlite3.dll``.  Same two tables and columns
(``SimulationRuns`` :63-68, ``RuntimeMetrics`` :112-114, with the key / timestamp columns of the commented-out
``EnsureTablesExist`` :7-44), same call surface (``SaveSimRunParams`` :46-96 -> run id, ``LogRuntimeMetrics`` :98-127),
written with Python's built-in ``sqlite3`` to any path -- or, with ``jsonl=True``, as one JSON object per line for hosts
without SQLite.  Host-side only: nothing here touches the solver.
"""
from __future__ import annotations

import json
import os
import sqlite3
import time

import numpy as np

_SCHEMA = """
PRAGMA foreign_keys = ON;
CREATE TABLE IF NOT EXISTS SimulationRuns (
    RunID INTEGER PRIMARY KEY AUTOINCREMENT,
    Size INTEGER, Diffusion REAL, Viscosity REAL, TimeStep REAL,
    SourceEnabled INTEGER, SourceStrength REAL, SourcePositionX REAL, SourcePositionY REAL,
    ObstacleEnabled INTEGER, ObstacleType TEXT, ObstaclePositionX REAL, ObstaclePositionY REAL,
    ObstacleRadius REAL, ObstacleWidth REAL, ObstacleHeight REAL,
    Timestamp DATETIME DEFAULT CURRENT_TIMESTAMP
);
CREATE TABLE IF NOT EXISTS RuntimeMetrics (
    MetricID INTEGER PRIMARY KEY AUTOINCREMENT,
    RunID INTEGER, Timestamp DATETIME DEFAULT CURRENT_TIMESTAMP,
    AverageDensity REAL, MaxVelocityMagnitude REAL, FrameRate REAL,
    FOREIGN KEY(RunID) REFERENCES SimulationRuns(RunID) ON DELETE CASCADE
);
"""
_RUN_COLUMNS = ("Size", "Diffusion", "Viscosity", "TimeStep", "SourceEnabled", "SourceStrength", "SourcePositionX",
                "SourcePositionY", "ObstacleEnabled", "ObstacleType", "ObstaclePositionX", "ObstaclePositionY",
                "ObstacleRadius", "ObstacleWidth", "ObstacleHeight")


class RunLog:
    def __init__(self, path: str, *, jsonl: bool = False, reference_quirks: bool = False):
        """reference_quirks=True reproduces SQL.cs:53-56 / :71: no run is recorded (run id -1, which also switches
        LogCurrentMetrics off, FluidSim.cs:580) when timeStep is the default 0.1f."""
        self.path, self.jsonl, self.quirks = path, jsonl, reference_quirks
        self._next_id = 1
        if jsonl:
            if os.path.exists(path):
                with open(path) as f:
                    self._next_id = 1 + sum(1 for line in f if '"table": "SimulationRuns"' in line)
        else:
            self.db = sqlite3.connect(path)
            self.db.executescript(_SCHEMA)

    def _emit(self, table, row):
        with open(self.path, "a") as f:
            f.write(json.dumps({"table": table, "timestamp": time.strftime("%Y-%m-%d %H:%M:%S"), **row}) + "\n")

    def save_sim_run_params(self, size, diffusion, viscosity, timeStep, sourceEnabled, sourceStrength, sourceX, sourceY,
                            obstacleEnabled, obstacleType, obstacleX, obstacleY, obstacleRadius, obstacleWidth,
                            obstacleHeight) -> int:
        """SQL.SaveSimRunParams (:46-96): returns the new run id, or -1."""
        if self.quirks and np.float32(timeStep) == np.float32(0.1):
            return -1
        values = (int(size), float(diffusion), float(viscosity), float(timeStep), int(bool(sourceEnabled)), float(sourceStrength),
                  float(sourceX), float(sourceY), int(bool(obstacleEnabled)), str(obstacleType), float(obstacleX), float(obstacleY),
                  float(obstacleRadius), float(obstacleWidth), float(obstacleHeight))
        if self.jsonl:
            run_id = self._next_id
            self._next_id += 1
            self._emit("SimulationRuns", {"RunID": run_id, **dict(zip(_RUN_COLUMNS, values))})
            return run_id
        cur = self.db.execute(f"INSERT INTO SimulationRuns ({', '.join(_RUN_COLUMNS)}) VALUES ({', '.join('?' * len(values))})", values)
        self.db.commit()
        return int(cur.lastrowid)

    def log_runtime_metrics(self, run_id, step, avg_density, max_velocity, frame_rate):
        """SQL.LogRuntimeMetrics (:98-127); `step` is accepted and ignored, as in the reference."""
        if run_id == -1:
            return
        row = {"RunID": int(run_id), "AverageDensity": float(avg_density), "MaxVelocityMagnitude": float(max_velocity),
               "FrameRate": float(frame_rate)}
        if self.jsonl:
            self._emit("RuntimeMetrics", row)
        else:
            self.db.execute("INSERT INTO RuntimeMetrics (RunID, AverageDensity, MaxVelocityMagnitude, FrameRate) VALUES (?, ?, ?, ?)",
                            tuple(row.values()))
            self.db.commit()

    def rows(self, table):
        if self.jsonl:
            with open(self.path) as f:
                return [r for r in map(json.loads, f) if r["table"] == table]
        cur = self.db.execute(f"SELECT * FROM {table}")
        names = [d[0] for d in cur.description]
        return [dict(zip(names, r)) for r in cur.fetchall()]

    def close(self):
        if not self.jsonl:
            self.db.close()
