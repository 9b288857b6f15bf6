"""realtime_scripts/active_microphones.py:4-45 -- unlike lib.directions, no +64 on unused_mics."""
import numpy as np

import realtime_scripts.config as config


def active_microphones():
    mode = config.mode
    rows = np.arange(0, config.rows, mode)
    columns = np.arange(0, config.columns * config.ACTIVE_ARRAYS, mode)
    per = config.rows * config.columns
    ids = np.arange(config.N_MICROPHONES, dtype=np.float64)
    mosaic = np.hstack([ids[a * per:(a + 1) * per].reshape(config.rows, config.columns)
                        for a in range(config.ACTIVE_ARRAYS)])
    try:
        unused = np.load('unused_mics.npy')
    except Exception:  # noqa: BLE001
        unused = []
    picked = [int(mosaic[r, c]) for r in rows for c in columns if mosaic[r, c] not in unused]
    return np.sort(picked)
