"""Result records of a simulation and their JSON / CSV wire format.

Wire-compatible with python_ldpc_app/results.py:10-117 (field names and order are
the format: ``results.json`` = ``{config{...}, snr_points[...], wall_clock_seconds,
adaptation_log}`` with ``indent=2, ensure_ascii=False``; ``results.csv`` = the 13
per-SNR fields in declaration order), so the reference's plot_results.py /
visualization.py keep working on files written here.
tests/golden/results_sample.{json,csv} were produced by the reference's writers.
"""
from __future__ import annotations

import csv
import dataclasses as dc
import json
from typing import List, Tuple


@dc.dataclass
class BlockResult:
    block_num: int
    snr_db: float
    decode_success: bool
    error_bits: int
    normalized_llr: float
    convergence_iteration: int      # pass index at which the syndrome vanished, -1 = never


@dc.dataclass
class SNRPointResult:
    snr_db: float
    ber: float
    fer: float
    avg_normalized_llr: float
    total_blocks: int
    successful_blocks: int
    failed_blocks: int
    avg_convergence_iterations: float
    matrix_path: str = ""
    modulation: int = 1
    max_iterations: int = 5
    interleaver: str = "none"
    encoding_method: str = "standard"


@dc.dataclass
class SimulationConfig:
    matrix_path: str
    n: int
    m: int
    k: int
    rate: float
    blocks: int
    max_iterations: int
    encoding_method: str
    interleaver_type: str
    decoder_type: str
    channel_mode: int
    modulation: int
    speed: float
    snr_range: Tuple[float, float, float]   # (start, end, step)
    threads: int
    timestamp: str
    interference_snr: float = 0.0
    p: float = 0.1


CSV_COLUMNS = [f.name for f in dc.fields(SNRPointResult)]


@dc.dataclass
class SimulationResult:
    config: SimulationConfig
    snr_points: List[SNRPointResult]
    wall_clock_seconds: float
    adaptation_log: List[dict] = dc.field(default_factory=list)

    def to_dict(self) -> dict:
        doc = dc.asdict(self)
        doc["config"]["snr_range"] = list(doc["config"]["snr_range"])
        return doc

    def to_json(self, filepath: str) -> None:
        with open(filepath, "w", encoding="utf-8") as fh:
            json.dump(self.to_dict(), fh, indent=2, ensure_ascii=False)

    def to_csv(self, filepath: str) -> None:
        if not self.snr_points:
            return
        with open(filepath, "w", newline="", encoding="utf-8") as fh:
            out = csv.DictWriter(fh, fieldnames=CSV_COLUMNS)
            out.writeheader()
            out.writerows({c: getattr(pt, c) for c in CSV_COLUMNS} for pt in self.snr_points)

    @classmethod
    def from_json(cls, filepath: str) -> "SimulationResult":
        with open(filepath, "r", encoding="utf-8") as fh:
            doc = json.load(fh)
        cfg = dict(doc["config"])
        cfg["snr_range"] = tuple(cfg["snr_range"])
        return cls(config=SimulationConfig(**cfg),
                   snr_points=[SNRPointResult(**pt) for pt in doc["snr_points"]],
                   wall_clock_seconds=doc["wall_clock_seconds"],
                   adaptation_log=doc.get("adaptation_log", []))
