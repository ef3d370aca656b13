"""Replicate drivers: --bootstrap and --windows runs of the reference are loops over independent
models (/root/reference/locator/locator.py:519-583 windows, :609-681 bootstrap).  Here the
independent models are work items of a dynamic queue served by one worker process per GPU
(``--gpus N``); there is no collective on the training path.  Everything that consumes numpy's
global random stream (bootstrap reseeds + site orders, per-window imputation / subsample / split)
is drawn serially in the parent, in the reference's order, and shipped with the work item, so the
indices do not depend on the number of GPUs.
"""
from __future__ import annotations

import copy
import os
import time
import traceback

import numpy as np


# ---------------------------------------------------------------------------------------------
# (de)serialisation of packed device matrices for the worker processes
# ---------------------------------------------------------------------------------------------
def _to_host(g):
    return {"words": g.words.cpu().numpy(), "n": g.n, "K": g.K}


def _to_dev(h):
    import torch
    from .genotypes import PackedGenotypes

    return PackedGenotypes(torch.as_tensor(h["words"]).cuda(), h["n"], h["K"])


def draw_bootstrap_orders(nsites, nboots):
    """The reference's draws for every replicate, in order (locator.py:637, :648-650):
    reseed from the stream, then resample the sites with replacement."""
    orders = []
    for _ in range(nboots):
        np.random.seed(np.random.choice(range(int(1e6)), 1))
        orders.append(np.random.choice(nsites, nsites, replace=True))
    return orders


def _run_item(L, item, base):
    """Train + predict one replicate inside the current process / on the current device."""
    args = L.args
    kind = item["kind"]
    if kind == "boot":
        traingen, testgen, predgen = base["traingen"], base["testgen"], base["predgen"]
        order = item["site_order"]
        tg, vg, pg = traingen.take_cols(order), testgen.take_cols(order), predgen.take_cols(order)
        L._seed_tag[0] = item["boot"] + 1
        print("starting bootstrap " + str(item["boot"]))
        L._run_one(tg, vg, base["trainlocs"], base["testlocs"], pg, base["norm"], base["pred"], base["samples"],
                   item["boot"], item["boot"])
    elif kind == "window":
        t1 = time.time()
        L._seed_tag[0] = item["index"] + 1
        original_out = args.out
        args.out = item["window_out"]  # predict_locs / history use the window-specific stem (locator.py:555-557)
        try:
            model = L.load_network(item["traingen"], args.dropout_prop)
            callbacks = L.load_callbacks(None)
            history, model = L.train_network(model, item["traingen"], item["testgen"], item["trainlocs"],
                                             item["testlocs"], callbacks)
            meanlong, sdlong, meanlat, sdlat = item["norm"]
            dists = L.predict_locs(model, item["predgen"], sdlong, meanlong, sdlat, meanlat, item["testlocs"],
                                   item["pred"], item["samples"], item["testgen"], history)
        finally:
            args.out = original_out
        if args.plot_history:
            L.plot_history(history, dists)
        print(f"Window run time {(time.time() - t1) / 60:.2f} minutes")
    else:
        raise ValueError(kind)


def _resolve(runner):
    """'package.module:function' -> callable (test hook: a CPU stand-in for _run_item)."""
    import importlib

    mod, fn = runner.split(":")
    return getattr(importlib.import_module(mod), fn)


def _worker(rank, n_gpus, args, base_host, task_q, result_q, runner=None):
    try:
        if runner is not None:  # host-logic tests: no CUDA, the runner consumes the raw item
            run = _resolve(runner)
            while True:
                item = task_q.get()
                if item is None:
                    break
                run(rank, args, item)
                result_q.put(("done", rank, item.get("boot", item.get("index"))))
            result_q.put(("exit", rank, None))
            return
        import torch

        torch.cuda.set_device(rank % max(1, torch.cuda.device_count()))
        from . import locator as L

        L.set_args(copy.copy(args))
        base = None
        if base_host is not None:
            base = dict(base_host)
            for k in ("traingen", "testgen", "predgen"):
                base[k] = _to_dev(base_host[k])
        while True:
            item = task_q.get()
            if item is None:
                break
            for k in ("traingen", "testgen", "predgen"):
                if k in item and isinstance(item[k], dict):
                    item[k] = _to_dev(item[k])
            _run_item(L, item, base)
            result_q.put(("done", rank, item.get("boot", item.get("index"))))
        result_q.put(("exit", rank, None))
    except Exception:  # surface the failure in the parent instead of hanging the queue
        result_q.put(("error", rank, traceback.format_exc()))


class ReplicatePool:
    """One worker process per GPU pulling work items from a shared queue."""

    def __init__(self, n_gpus, args, base=None, runner=None):
        import multiprocessing as mp

        ctx = mp.get_context("spawn")
        self.n = n_gpus
        self.task_q = ctx.Queue()
        self.result_q = ctx.Queue()
        base_host = None
        if base is not None:
            base_host = dict(base)
            for k in ("traingen", "testgen", "predgen"):
                base_host[k] = _to_host(base[k])
        self.procs = [ctx.Process(target=_worker, args=(r, n_gpus, args, base_host, self.task_q, self.result_q, runner))
                      for r in range(n_gpus)]
        for p in self.procs:
            p.start()
        self.submitted = 0

    def submit(self, item):
        item = dict(item)
        for k in ("traingen", "testgen", "predgen"):
            if k in item and not isinstance(item[k], dict):
                item[k] = _to_host(item[k])
        self.task_q.put(item)
        self.submitted += 1

    def close(self):
        for _ in self.procs:
            self.task_q.put(None)
        done = exited = 0
        errors = []
        while exited < self.n and not errors:
            kind, rank, payload = self.result_q.get()
            if kind == "done":
                done += 1
            elif kind == "exit":
                exited += 1
            else:
                errors.append((rank, payload))
        for p in self.procs:
            if errors:
                p.terminate()
            p.join()
        if errors:
            raise RuntimeError(f"replicate worker {errors[0][0]} failed:\n{errors[0][1]}")
        if done != self.submitted:
            raise RuntimeError(f"{self.submitted - done} replicates were not completed")


# ---------------------------------------------------------------------------------------------
# drivers called from locator.main()
# ---------------------------------------------------------------------------------------------
def run_bootstrap(L, traingen, testgen, trainlocs, testlocs, predgen, norm, pred, samples):
    args = L.args
    # 1. initial full run (locator.py:611-632)
    L._seed_tag[0] = 0
    L._run_one(traingen, testgen, trainlocs, testlocs, predgen, norm, pred, samples, "FULL", "FULL")
    # 2. every replicate's draws, serially, as the reference's loop would take them
    orders = draw_bootstrap_orders(traingen.K, args.nboots)
    base = {"traingen": traingen, "testgen": testgen, "predgen": predgen, "trainlocs": trainlocs,
            "testlocs": testlocs, "norm": norm, "pred": pred, "samples": samples}
    n_gpus = max(1, int(getattr(args, "gpus", 1) or 1))
    if n_gpus == 1:
        for boot, order in enumerate(orders):
            _run_item(L, {"kind": "boot", "boot": boot, "site_order": order}, base)
        return
    pool = ReplicatePool(n_gpus, args, base)
    for boot, order in enumerate(orders):
        pool.submit({"kind": "boot", "boot": boot, "site_order": order})
    pool.close()


def window_bounds(positions, start, stop, size):
    """(i, a, b) per window; the slice is gt[a:b] -- SNP b itself is excluded, as in locator.py:531-538."""
    positions = np.asarray(positions)
    for i in np.arange(start, stop, size):
        mask = np.logical_and(positions >= i, positions < i + size)
        w = np.argwhere(mask)
        yield int(i), int(np.min(w)), int(np.max(w))


def run_windows(L, genotypes, samples):
    args = L.args
    positions = np.array(genotypes.positions)
    start = int(args.window_start)
    stop = np.max(positions) if args.window_stop == None else int(args.window_stop)  # noqa: E711
    size = int(float(args.window_size))
    n_gpus = max(1, int(getattr(args, "gpus", 1) or 1))
    pool = ReplicatePool(n_gpus, args) if n_gpus > 1 else None
    for index, (i, a, b) in enumerate(window_bounds(positions, start, stop, size)):
        print(f"\nProcessing window {i}-{i+size}")
        print(f"SNPs {a}-{b}")
        sub = genotypes[a:b]
        sample_data, locs = L.sort_samples(samples, sub)
        meanlong, sdlong, meanlat, sdlat, locs = L.normalize_locs(locs)
        ac = L.filter_snps(sub)
        train, test, traingen, testgen, trainlocs, testlocs, pred, predgen = L.split_train_test(ac, locs)
        item = {"kind": "window", "index": index, "window_out": f"{args.out}_{i}-{i+size-1}", "traingen": traingen,
                "testgen": testgen, "predgen": predgen, "trainlocs": trainlocs, "testlocs": testlocs,
                "norm": (meanlong, sdlong, meanlat, sdlat), "pred": pred, "samples": samples}
        if pool is None:
            _run_item(L, item, None)
        else:
            pool.submit(item)
    if pool is not None:
        pool.close()
