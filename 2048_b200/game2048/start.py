"""Bootstrap of the drop-in package (reference: game2048/start.py).

The reference's start.py wires S3 credentials, Dash timers and a cloud logger; all of that is OUT OF SCOPE
(SURVEY.md section 2).  What the hot path and its callers need from it is kept with the same names:
  * the names the star-import chain re-exports (np, random, pickle, time, deque, Thread, ...), start.py:1-17
  * the cooperative-stop globals GAME_PANE / AGENT_PANE / RUNNING and dash_intervals, start.py:23-32
  * load_s3 / save_s3 / list_names_s3 / delete_s3 / is_data_there and Logger, start.py:67-158 -- here they
    store the same object names as FILES under a local directory ($B2048_STORAGE, default ./b2048_storage),
    so the two-object agent layout ('a/<name>.pkl' + 'weights/<name>.pkl', r_learning.py:166-200) round-trips
    without a network.
"""
from datetime import datetime, timedelta  # noqa: F401
import json
import os
import pickle
import random  # noqa: F401
import sys  # noqa: F401
import time  # noqa: F401
from collections import deque  # noqa: F401
from threading import Thread  # noqa: F401

import numpy as np  # noqa: F401

LOCAL = os.environ.get("S3_URL", "local")
dash_intervals = {"refresh_sec": 60, "vc_sec": 300, "initiate_logs": 3000, "logs": 1000}
dash_intervals["refresh"] = dash_intervals["refresh_sec"] * 1000
dash_intervals["check_run"] = dash_intervals["refresh_sec"] * 2
dash_intervals["vc"] = dash_intervals["vc_sec"] * 1000
dash_intervals["next"] = dash_intervals["refresh_sec"] + 180
LOWEST_SPEED = 50

GAME_PANE = {}
AGENT_PANE = {}
RUNNING = {}


def storage_dir():
    d = os.environ.get("B2048_STORAGE", os.path.join(os.getcwd(), "b2048_storage"))
    os.makedirs(d, exist_ok=True)
    return d


def _path(name):
    """object name -> file under storage_dir(); names that would escape it ('..', absolute paths) are refused"""
    root = os.path.realpath(storage_dir())
    p = os.path.realpath(os.path.join(root, *str(name).split("/")))
    if os.path.commonpath([root, p]) != root or p == root:
        raise ValueError(f"object name {name!r} leaves the storage directory")
    os.makedirs(os.path.dirname(p), exist_ok=True)
    return p


def list_names_s3():
    root = storage_dir()
    out = []
    for base, _, files in os.walk(root):
        for f in files:
            out.append(os.path.relpath(os.path.join(base, f), root).replace(os.sep, "/"))
    return sorted(out)


def is_data_there(name):
    return bool(name) and os.path.isfile(_path(name))


def delete_s3(name):
    if is_data_there(name):
        os.remove(_path(name))


def load_s3(name):
    """json / txt / pkl object by name, or None (start.py:84-101)"""
    if not name or not is_data_there(name):
        return None
    ext = name.rsplit(".", 1)[-1]
    if ext == "json":
        with open(_path(name), "r", encoding="utf-8") as f:
            return json.load(f)
    if ext == "txt":
        with open(_path(name), "r") as f:
            return f.read()
    if ext == "pkl":
        with open(_path(name), "rb") as f:
            return pickle.load(f)
    return None


def save_s3(data, name):
    """start.py:104-119; returns 1 on success, 0 for an unknown extension"""
    ext = name.rsplit(".", 1)[-1]
    if ext == "json":
        with open(_path(name), "w") as f:
            json.dump(data, f)
    elif ext == "txt":
        with open(_path(name), "w") as f:
            f.write(data)
    elif ext == "pkl":
        with open(_path(name), "wb") as f:
            pickle.dump(data, f, -1)
    else:
        return 0
    return 1


class Logger:
    """start.py:144-158: append-only text log addressed by object name"""
    msg = {"welcome": "Welcome! Let's do something interesting. Choose MODE of action!",
           "collapse": "Current process collapsed!"}

    def __init__(self, log_file):
        self.file = log_file
        if not is_data_there(self.file):
            save_s3("", self.file)

    def add(self, text):
        if text:
            with open(_path(self.file), "a") as f:
                f.write("\n" + str(text))
