"""filter_obs: drop-in for oneka/utilities.py:34-144 (one-off preprocessing, host side)."""
import logging

import numpy as np

log = logging.getLogger("Oneka")


def filter_obs(observations, wellfield, buffer):
    """Drop observations within `buffer` of any well, then merge runs of consecutive
    observations closer than 1 m into their minimum-variance average (weights 1/sigma^2).

    Same rules and same order of operations as the reference: the distance test is
    hypot(..) <= buffer (oneka/utilities.py:97-103); duplicates are detected only among
    CONSECUTIVE retained observations, comparing each with the first of its run (:110-113)."""
    TOO_CLOSE = 1.0
    wxy = np.array([[w[0], w[1]] for w in wellfield], dtype=float).reshape(-1, 2)
    obs = []
    for ob in observations:
        if len(wxy) and np.any(np.hypot(ob[0] - wxy[:, 0], ob[1] - wxy[:, 1]) <= buffer):
            log.info('observation removed: %s is too close to a well', (ob,))
            continue
        obs.append(ob)
    retained = []
    i = 0
    while i < len(obs):
        j = i + 1
        while (j < len(obs)) and (np.hypot(obs[i][0] - obs[j][0], obs[i][1] - obs[j][1]) < TOO_CLOSE):
            j += 1
        if j - i > 1:
            num = 0
            den = 0
            for k in range(i, j):
                num += obs[k][2] / obs[k][3] ** 2
                den += 1 / obs[k][3] ** 2
            retained.append((obs[i][0], obs[i][1], num / den, np.sqrt(1 / den)))
        else:
            retained.append(obs[i])
        i = j
    log.info('active observations: %d', len(retained))
    return retained
