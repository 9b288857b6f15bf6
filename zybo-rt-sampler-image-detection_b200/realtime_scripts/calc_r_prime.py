"""realtime_scripts/calc_r_prime.py:9-24 -- differs from lib.directions.calc_r_prime: array
separation terms and a 0.11 m camera offset on y."""
import numpy as np

import realtime_scripts.active_microphones as am
import realtime_scripts.config as config

camera_offset = 0.11      # [m]


def calc_r_prime(d):
    half = d / 2
    pos = np.zeros((2, config.N_MICROPHONES))
    k = 0
    for a in range(config.ACTIVE_ARRAYS):
        a = -a
        for row in range(config.rows):
            for col in range(config.columns):
                pos[0, k] = (-col * d - half + a * config.columns * d + a * config.ARRAY_SEPARATION
                             + config.columns * config.ACTIVE_ARRAYS * half)
                pos[1, k] = row * d - config.rows * half + half - camera_offset
                k += 1
    pos[0, :] += (config.ACTIVE_ARRAYS - 1) * config.ARRAY_SEPARATION / 2
    active = am.active_microphones()
    return pos, pos[:, active]
