"""args.json persistence: mirror of logger_utils.py:8-20 (the TensorBoard summaries of :23-33 are opt-in
observability, out of scope for the hot path -- global_norm is still produced by the optimizer kernel)."""
import json
import os


def load_args(path):
    if path is None:
        return {}
    with open(path, 'r') as f:
        return json.load(f)


def save_args(args, folder, file_name='args.json'):
    args = vars(args)
    if not os.path.exists(folder):
        os.makedirs(folder)
    with open(os.path.join(folder, file_name), 'w') as f:
        return json.dump(args, f)
