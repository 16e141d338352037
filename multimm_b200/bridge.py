"""SimulationEngine — embedding API, mirror of the reference's ``bridge.py`` (src/multimm/bridge.py:16-164):
schema export, parameter validation, in-process run (3 attempts, bridge.py:102-118) and subprocess
run.  Differences, both forced by the engine: a device failure (``multimm_b200.Error`` whose text
contains "CUDA error", the substring bridge.py:70-75 looks for) is retried but never re-run on a
CPU platform — there is none; and ensemble members of one call are dealt to the visible GPUs.
"""
from __future__ import annotations

import contextlib
import logging
import os
import subprocess
import sys
from typing import Any, Dict

from ._lib import Error
from .config import SimulationConfig
from .run import run_replica, visible_devices, write_config

logger = logging.getLogger("multimm_b200")

PLATFORM_ERRORS = ("Error initializing context", "CUDA error")
ATTEMPTS = 3  # bridge.py:102-118


def _metadata_dir(config) -> str:
    meta = os.path.join(config.OUT_PATH, "metadata")
    os.makedirs(meta, exist_ok=True)
    return meta


@contextlib.contextmanager
def _run_log(path: str):
    """The package logger also writes to <OUT_PATH>/metadata/output.log for the duration of a run."""
    handler = logging.FileHandler(path, mode="w")
    handler.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
    level = logger.level
    logger.addHandler(handler)
    if level == logging.NOTSET or level > logging.INFO:
        logger.setLevel(logging.INFO)
    try:
        yield
    finally:
        logger.removeHandler(handler)
        handler.close()
        logger.setLevel(level)


def _with_retries(job, device: int):
    """An engine failure is logged and the job repeated; the last failure propagates."""
    for k in range(1, ATTEMPTS + 1):
        try:
            return job()
        except Error as e:
            kind = "platform" if any(s in str(e) for s in PLATFORM_ERRORS) else "engine"
            logger.error(f"Simulation failed ({kind} error) on device {device}: {e}")
            if k == ATTEMPTS:
                raise
            logger.warning(f"Attempt {k} failed, retrying...")


class SimulationEngine:
    @classmethod
    def get_schema(cls) -> Dict[str, Any]:
        return SimulationConfig.model_json_schema()

    @classmethod
    def validate_params(cls, params: Dict[str, Any]) -> Dict[str, Any]:
        return SimulationConfig(**params).model_dump()

    @classmethod
    def run_in_process(cls, config_params: Dict[str, Any], fallback_to_cpu: bool = False) -> str:
        """Runs synchronously in this process and returns the path of ``config_auto.ini``.
        ``fallback_to_cpu`` is accepted for signature compatibility; True raises, because silently
        running somewhere else is exactly what this engine does not do."""
        if fallback_to_cpu:
            raise Error(-3, "fallback_to_cpu=True: this engine has no CPU platform to fall back to")
        config = SimulationConfig(**config_params)
        meta = _metadata_dir(config)
        params = config.model_dump()
        devices = visible_devices(config)
        with _run_log(os.path.join(meta, "output.log")):
            write_config(config)
            if config.GENERATE_ENSEMBLE and config.N_ENSEMBLE is not None:
                # bridge.py:96-100: member i runs with seed start + i into <OUT_PATH>_<i+1>
                jobs = [(int(config.SHUFFLING_SEED) + i, f"{config.OUT_PATH}_{i + 1}", devices[i % len(devices)])
                        for i in range(config.N_ENSEMBLE)]
            else:
                jobs = [(int(config.SHUFFLING_SEED), config.OUT_PATH, devices[0])]
            for seed, out_path, device in jobs:
                _with_retries(lambda: run_replica(params, seed, out_path, device, archive=False), device)
        return os.path.join(meta, "config_auto.ini")

    @classmethod
    def run_subprocess(cls, config_params: Dict[str, Any]) -> str:
        config = SimulationConfig(**config_params)
        meta = _metadata_dir(config)
        config_path = write_config(config)
        with open(os.path.join(meta, "output.log"), "w") as log_file:
            subprocess.run([sys.executable, "-m", "multimm_b200.run", "-c", config_path], stdout=log_file,
                           stderr=subprocess.STDOUT, text=True, check=True)
        return config_path
