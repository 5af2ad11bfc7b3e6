"""Scripted scenarios for the queue-connected pipeline modules (pipeline/pipeline_module.py, datasets/data_module.py,
slam/slam_module.py, fusion/fusion_module.py).  `run(classes)` drives a set of module classes — the reference's own
(make_golden_pipeline.py) or this repo's (tests/test_cpu_shim.py) — with scripted workers and returns a JSON-able log:
every spin()'s return value, what arrived on the queues / callbacks, shutdown flags and failure notifications."""
import queue
import types


class Dataset:
    def __init__(self, n):
        self.items = [{"k": [i], "payload": i * i} for i in range(n)]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


class Slam:
    """returns [state, viz] per input; `falsy_at`: outputs that are falsy (module must shut down, slam_module.py:12-14);
    `none_at`: None outputs (nothing pushed, failure callbacks fire); stop_condition() turns true after `stop_after` calls"""

    def __init__(self, log, falsy_at=(), none_at=(), stop_after=10 ** 9):
        self.log, self.falsy_at, self.none_at, self.stop_after, self.n = log, set(falsy_at), set(none_at), stop_after, 0

    def __call__(self, inp):
        self.log.append(["slam_in", sorted(inp), inp["data"]["k"][0] if "data" in inp else None])
        self.n += 1
        if self.n - 1 in self.none_at:
            return None
        if self.n - 1 in self.falsy_at:
            return False
        return ["state%d" % self.n, {"kf_idx": self.n}]

    def stop_condition(self):
        return self.n >= self.stop_after


class Fusion:
    def __init__(self, log, stop_after_idle=3, outputs=True):
        self.log, self.stop_after_idle, self.idle, self.outputs = log, stop_after_idle, 0, outputs

    def fuse(self, packets):
        if packets:
            v = packets.get("slam")             # a falsy SLAM output is pushed like any other (pipeline_module.py:103-106)
            self.log.append(["fuse", sorted(packets), v[1]["kf_idx"] if v else repr(v)])
        else:
            self.idle += 1
            self.log.append(["fuse_idle", bool(packets), packets is False])
        return {"mesh": self.idle} if self.outputs else None

    def stop_condition(self):
        return self.idle >= self.stop_after_idle


def _args(parallel):
    return types.SimpleNamespace(parallel_run=parallel)


def _ready(mod, attr, worker):
    setattr(mod, attr, worker)
    mod.is_initialized = True               # the real initialize_module imports the GPU workers; they are scripted here
    return mod


def sequential(classes, n_frames=5, falsy_at=(), none_at=(), stop_after=10 ** 9):
    """examples/slam_demo.py:160-181 in sequential mode"""
    DataModule, SlamModule, FusionModule = classes
    log = []
    data = _ready(DataModule("nerf", _args(False)), "dataset", Dataset(n_frames))
    slam = _ready(SlamModule("VioSLAM", _args(False), device="cpu"), "slam", Slam(log, falsy_at, none_at, stop_after))
    fus = _ready(FusionModule("nerf", _args(False), device="cpu"), "fusion", Fusion(log))
    dq, sq, gq = queue.Queue(), queue.Queue(), queue.Queue()
    fails = {"slam": 0, "data": 0}
    data.register_output_queue(dq); slam.register_input_queue("data", dq)
    slam.register_output_queue(sq); fus.register_input_queue("slam", sq)
    fus.register_output_queue(gq)
    cb_seen = []
    slam.register_output_callback(lambda p: cb_seen.append(p[0]))
    slam.register_on_failure_callback(lambda: fails.__setitem__("slam", fails["slam"] + 1))
    data.register_on_failure_callback(lambda: fails.__setitem__("data", fails["data"] + 1))
    steps = []
    for _ in range(40):
        r = [bool(data.spin())]
        r.append(bool(slam.spin()) if r[0] else None)
        r.append(bool(fus.spin()) if r[0] and r[1] else None)
        steps.append(r + [bool(data.shutdown), bool(slam.shutdown), bool(fus.shutdown), dq.qsize(), sq.qsize(), gq.qsize()])
        if not all(x for x in r):
            break
    return {"log": log, "steps": steps, "callbacks": cb_seen, "fails": fails}


def parallel_loops(classes):
    """parallel_run=True: spin() loops until the module shuts itself down; queues pre-filled, one module at a time"""
    DataModule, SlamModule, FusionModule = classes
    log = []
    data = _ready(DataModule("nerf", _args(True)), "dataset", Dataset(4))
    dq = queue.Queue()
    data.register_output_queue(dq)
    r_data = data.spin()
    n_data = dq.qsize()
    slam = _ready(SlamModule("VioSLAM", _args(True), device="cpu"), "slam", Slam(log, stop_after=3))
    sq = queue.Queue()
    slam.register_input_queue("data", dq); slam.register_output_queue(sq)
    r_slam = slam.spin()
    fus = _ready(FusionModule("nerf", _args(True), device="cpu"), "fusion", Fusion(log, stop_after_idle=2))
    gq = queue.Queue()
    fus.register_input_queue("slam", sq); fus.register_output_queue(gq)
    r_fus = fus.spin()
    return {"log": log, "returns": [bool(r_data), bool(r_slam), bool(r_fus)], "sizes": [n_data, dq.qsize(), sq.qsize(), gq.qsize()],
            "flags": [bool(data.shutdown), bool(slam.shutdown), bool(fus.shutdown), int(data.idx)]}


def restart_and_failures(classes):
    """shutdown_module / restart, a push into a consumer that raises, an unknown dataset / slam / fusion name"""
    DataModule, SlamModule, FusionModule = classes
    out = {}
    data = _ready(DataModule("nerf", _args(False)), "dataset", Dataset(2))

    class Bad:
        def put(self, p):
            raise RuntimeError("closed")
    good = queue.Queue()
    data.register_output_queue(Bad()); data.register_output_queue(good)
    out["spin_with_bad_consumer"] = [bool(data.spin()), good.qsize(), bool(data.shutdown)]
    data.shutdown_module()
    out["after_shutdown"] = [bool(data.shutdown), bool(data.spin())]
    data.restart()
    out["after_restart"] = [bool(data.shutdown), bool(data.spin()), good.qsize(), int(data.idx)]
    errs = []
    for cls, name in ((DataModule, "no_such_dataset"), (SlamModule, "OtherSLAM"), (FusionModule, "no_such_fusion")):
        m = cls(name, _args(False)) if cls is DataModule else cls(name, _args(False), device="cpu")
        try:
            m.initialize_module()
            errs.append("no error")
        except NotImplementedError:
            errs.append("NotImplementedError")
        except Exception:                       # noqa: BLE001
            errs.append("Exception")
    out["unknown_names"] = errs
    empty = SlamModule("VioSLAM", _args(False), device="cpu")
    empty = _ready(empty, "slam", Slam([]))
    empty.register_input_queue("data", queue.Queue())
    out["empty_input"] = [empty.get_input_packet() is None, bool(empty.spin())]
    f = _ready(FusionModule("nerf", _args(False), device="cpu"), "fusion", Fusion([]))
    f.register_input_queue("slam", queue.Queue())
    out["fusion_empty_input_is_false"] = f.get_input_packet() is False
    return out


def run(classes):
    return {"sequential": sequential(classes),
            "sequential_falsy": sequential(classes, falsy_at=(2,)),
            "sequential_none": sequential(classes, none_at=(1,)),
            "sequential_stop": sequential(classes, n_frames=6, stop_after=3),
            "parallel": parallel_loops(classes),
            "misc": restart_and_failures(classes)}
