"""Host-side helpers with the reference's names (rllab/sampler/parallel_sampler.py).  The process pool itself
(initialize / populate_task / sample_paths) has no counterpart: LaneSampler replaces it with one CUDA launch."""
import numpy as np


def _truncate_tensor_dict(d, n):
    out = dict()
    for k, v in d.items():
        out[k] = _truncate_tensor_dict(v, n) if isinstance(v, dict) else v[:n]
    return out


def truncate_paths(paths, max_samples):
    """rllab/sampler/parallel_sampler.py:129-155, verbatim semantics: drop extra paths from the end, then shorten the
    last one so that the total number of samples is exactly max_samples.  Works on the list-of-path-dicts wire format
    (e.g. `samples_data["paths"]` / `LanePaths.to_paths()`); does not modify its input."""
    paths = list(paths)
    total_n_samples = sum(len(path["rewards"]) for path in paths)
    while len(paths) > 0 and total_n_samples - len(paths[-1]["rewards"]) >= max_samples:
        total_n_samples -= len(paths.pop(-1)["rewards"])
    if len(paths) > 0:
        last_path = paths.pop(-1)
        truncated_last_path = dict()
        truncated_len = len(last_path["rewards"]) - (total_n_samples - max_samples)
        for k, v in last_path.items():
            if k in ["observations", "actions", "rewards", "advantages", "returns"]:
                truncated_last_path[k] = np.asarray(v)[:truncated_len]
            elif k in ["env_infos", "agent_infos"]:
                truncated_last_path[k] = _truncate_tensor_dict(v, truncated_len)
            else:
                raise NotImplementedError
        paths.append(truncated_last_path)
    return paths
