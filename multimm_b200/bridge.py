"""SimulationEngine — embedding API, mirror of the reference's ``bridge.py`` (src/multimm/bridge.py:16-164):
schema export, parameter validation, in-process run (3 attempts, bridge.py:102-118) and subprocess
run.  Differences, both forced by the engine: a device failure (``multimm_b200.Error`` whose text
contains "CUDA error", the substring bridge.py:70-75 looks for) is retried but never re-run on a
CPU platform — there is none; and ensemble members of one call are dealt to the visible GPUs.
"""
from __future__ import annotations

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
        os.makedirs(config.OUT_PATH, exist_ok=True)
        meta = os.path.join(config.OUT_PATH, "metadata")
        os.makedirs(meta, exist_ok=True)
        handler = logging.FileHandler(os.path.join(meta, "output.log"), mode="w")
        handler.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        logger.addHandler(handler)
        old_level = logger.level
        if old_level == logging.NOTSET or old_level > logging.INFO:
            logger.setLevel(logging.INFO)

        def attempt(params: dict, seed: int, out_path: str, device: int):
            for k in range(3):
                try:
                    return run_replica(params, seed, out_path, device, archive=False)
                except Error as e:
                    kind = "platform" if any(s in str(e) for s in PLATFORM_ERRORS) else "engine"
                    logger.error(f"Simulation failed ({kind} error) on device {device}: {e}")
                    if k == 2:
                        raise
                    logger.warning(f"Attempt {k + 1} failed, retrying...")

        try:
            write_config(config)
            params = config.model_dump()
            devices = visible_devices(config)
            if config.GENERATE_ENSEMBLE and config.N_ENSEMBLE is not None:
                base, seed0 = config.OUT_PATH, int(config.SHUFFLING_SEED)
                for i in range(config.N_ENSEMBLE):  # bridge.py:96-100: seeds start+i, paths <base>_<i+1>
                    attempt(params, seed0 + i, f"{base}_{i + 1}", devices[i % len(devices)])
            else:
                attempt(params, int(config.SHUFFLING_SEED), config.OUT_PATH, devices[0])
        finally:
            logger.removeHandler(handler)
            handler.close()
            logger.setLevel(old_level)
        return os.path.join(meta, "config_auto.ini")

    @classmethod
    def run_subprocess(cls, config_params: Dict[str, Any]) -> str:
        config = SimulationConfig(**config_params)
        meta = os.path.join(config.OUT_PATH, "metadata")
        os.makedirs(meta, exist_ok=True)
        config_path = write_config(config)
        with open(os.path.join(meta, "output.log"), "w") as log_file:
            subprocess.run([sys.executable, "-m", "multimm_b200.run", "-c", config_path], stdout=log_file,
                           stderr=subprocess.STDOUT, text=True, check=True)
        return config_path
