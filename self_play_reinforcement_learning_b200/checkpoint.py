"""On-disk formats of the reference, so existing runs can be resumed and produced.

* model checkpoints: ``torch.save({"model": state_dict}, <save_dir>/<start_time>/model-<iso timestamp>:<games>)``
  (updateworker.py:111-117, name built at self_play_parallel.py:263-267)
* replay memory: ``pickle.dump(Memory)`` to ``memory-<iso timestamp>:<size>`` with the previous file removed
  (updateworker.py:119-139)
* discovery: lexicographically newest non-empty run folder, then the lexicographically newest file with the prefix
  (base_worker.py:44-62)
Checkpoints feed the engine through ``nets.pack_tower_blob(module)`` / ``BatchedSelfPlay.load_weights``.
"""
import datetime
import os
import pickle

import torch


def model_file_name(save_dir, start_time, games_played, now=None):
    now = now or datetime.datetime.now()
    return os.path.join(save_dir, start_time, "model-" + now.isoformat() + ":" + str(games_played))


def save_model(network, saved_name):
    os.makedirs(os.path.dirname(saved_name), exist_ok=True)
    torch.save({"model": network.state_dict()}, saved_name)
    return saved_name


def load_model(network, model_file, map_location="cpu"):
    checkpoint = torch.load(model_file, map_location=map_location)
    network.load_state_dict(checkpoint["model"])
    return network


def save_memory(memory, save_dir, start_time, previous=None, now=None):
    now = now or datetime.datetime.now()
    name = os.path.join(save_dir, start_time, "memory-" + now.isoformat() + ":" + str(len(memory)))
    os.makedirs(os.path.dirname(name), exist_ok=True)
    with open(name, "wb") as f:
        pickle.dump(memory, f)
    if previous and os.path.exists(previous):
        os.remove(previous)
    return name


def load_memory(memory_file):
    with open(memory_file, "rb") as f:
        return pickle.load(f)


def recent_save_file(save_dir, start_time=None, prev_run=False, starting_str="model"):
    """base_worker.py:44-62: newest (by name) non-empty run folder -- excluding ``start_time`` when ``prev_run`` -- and
    in it the newest (by name) file whose name starts with ``starting_str``."""
    folders = [os.path.join(save_dir, f) for f in os.listdir(save_dir)
               if not os.path.isfile(os.path.join(save_dir, f)) and not (prev_run and f == start_time)]
    non_empty = [f for f in folders if os.listdir(f)]
    recent_folder = max(non_empty)
    saves = [os.path.join(recent_folder, f) for f in os.listdir(recent_folder)
             if os.path.isfile(os.path.join(recent_folder, f)) and f.startswith(starting_str)]
    return max(saves)
