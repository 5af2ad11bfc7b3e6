"""Queue-connected pipeline modules with the reference's surface (pipeline/pipeline_module.py `MIMOPipelineModule`,
datasets/data_module.py `DataModule`, slam/slam_module.py `SlamModule`, fusion/fusion_module.py `FusionModule`,
slam/vio_slam.py `VioSLAM`) so that the reference's entry point examples/slam_demo.py runs unchanged when
`nerf_slam_b200/shim` is first on sys.path.  Control plane only — nothing of the hot path lives here; gtsam is not needed
(the reference's back-end receives empty values / factors from the front end and is re-created after every step,
slam/meta_slam.py:44, visual_frontend.py:248-249).

One generic module class does the queue work; the three concrete modules differ only in how they build their worker
(lazily, inside the process that spins them — the reason the reference has `initialize_module`) and in what one step
does with an input."""
import logging
import queue as _queue

import numpy as np

_log = logging.getLogger("nerf_slam.pipeline")


class MIMOPipelineModule:
    """many-in / many-out module: named input queues, output queues + callbacks, `spin()` = one step in sequential mode
    (returns True while alive) or a loop until shutdown in parallel mode (returns False at the end)"""

    def __init__(self, name_id, parallel_run, args=None, grad=False):
        self.name, self.parallel_run, self.args, self.grad = name_id, bool(parallel_run), args, grad
        self.shutdown = self.is_initialized = self.is_thread_working = False
        self.input_queues, self.output_queues, self.output_callbacks, self.on_failure_callbacks = {}, [], [], []

    # -- wiring
    def register_input_queue(self, name, q):
        self.input_queues[name] = q

    def register_output_queue(self, q):
        self.output_queues.append(q)

    def register_output_callback(self, cb):
        self.output_callbacks.append(cb)

    def register_on_failure_callback(self, cb):
        self.on_failure_callbacks.append(cb)

    # -- life cycle
    def initialize_module(self):
        self.is_initialized = True
        return True

    def shutdown_queues(self):
        pass

    def shutdown_module(self):
        self.shutdown_queues()
        self.shutdown = True

    def restart(self):
        self.shutdown = False

    # -- data movement
    def get_input_packet(self, timeout=0.1):
        got = {}
        for name, q in self.input_queues.items():
            try:
                got[name] = q.get(timeout=timeout) if self.parallel_run else q.get_nowait()
            except _queue.Empty:
                pass
            except Exception as e:            # noqa: BLE001 (torch.multiprocessing queues raise their own Empty)
                _log.debug(e)
        return got or None

    def push_output_packet(self, packet):
        ok = True
        for sink in [q.put for q in self.output_queues] + self.output_callbacks:
            try:
                sink(packet)
            except Exception as e:            # noqa: BLE001 (a closed consumer must not kill the producer)
                _log.warning(e); ok = False
        return ok

    def spin_once(self, inp):
        raise NotImplementedError

    def spin(self):
        if not self.is_initialized:
            self.initialize_module()
        while not self.shutdown:
            inp = self.get_input_packet()
            self.is_thread_working = True
            if inp is not None:
                out = self.spin_once(inp)
                if out is None:
                    for cb in self.on_failure_callbacks:
                        cb()
                elif not self.push_output_packet(out):
                    _log.warning("Module %s: output push failed", self.name)
            self.is_thread_working = False
            if not self.parallel_run:
                return True
        return False


class _WorkerModule(MIMOPipelineModule):
    """module around a lazily built worker object"""
    attr = "worker"

    def __init__(self, name, args, device="cpu"):
        super().__init__(name, getattr(args, "parallel_run", False), args)
        self.device = device

    def build(self):
        raise NotImplementedError

    def initialize_module(self):
        setattr(self, self.attr, self.build())
        return super().initialize_module()


class DataModule(_WorkerModule):
    """dataset reader as a source module: emits dataset[idx] per spin, shuts itself down at the end"""
    attr = "dataset"

    def __init__(self, name, args, device="cpu"):
        super().__init__(name, args, device)
        self.idx = -1

    def build(self):
        if self.name not in ("nerf", "replica"):
            raise Exception(f"dataset format '{self.name}' is outside the hot-path scope (DESIGN.md): use 'nerf'")
        from .datasets import NeRFDataset
        return NeRFDataset(self.args, self.device)

    def get_input_packet(self, timeout=0.0):
        return True

    def spin_once(self, _):
        self.idx += 1
        if self.idx < len(self.dataset):
            return self.dataset[self.idx]
        print("Stopping data module!")
        self.shutdown_module()
        return None


class SLAM:
    """front end -> (no-op) back end, slam/meta_slam.py:26-47"""

    def __init__(self, name, args, device):
        self.name, self.args, self.device = name, args, device
        self.state = self.delta = None

    def forward(self, batch):
        assert "data" in batch
        out = self._frontend(batch["data"], self.state, self.delta)
        if out is False:
            return out
        x0, factors, viz_out = out
        self.state, self.delta = self._backend(factors, x0)
        return [self.state, viz_out]

    __call__ = forward


# the reference's hard-coded first pose (slam/vio_slam.py:91-96: a Replica camera pose; data, not code)
WORLD_T_IMU_T0 = np.array([[-7.6942980e-02, -3.1037781e-01, 9.4749427e-01, 8.9643948e-02],
                           [-2.8366595e-10, -9.5031142e-01, -3.1130061e-01, 4.1829333e-01],
                           [9.9703550e-01, -2.3952398e-02, 7.3119797e-02, 4.8306200e-01],
                           [0.0, 0.0, 0.0, 1.0]])


class VioSLAM(SLAM):
    """slam/vio_slam.py:78-127 without gtsam: imu_T_cam0 = identity (:88); `args.world_T_imu_t0` (the commented-out
    intent of :90) overrides the hard-coded first pose when given"""

    def __init__(self, name, args, device):
        super().__init__(name, args, device)
        from .frontend import RaftVisualFrontend
        w = getattr(args, "world_T_imu_t0", None)
        self.visual_frontend = RaftVisualFrontend(WORLD_T_IMU_T0 if w is None else np.asarray(w, np.float64), np.eye(4),
                                                  args, device=device)
        self.last_state = None

    def stop_condition(self):
        return self.visual_frontend.stop_condition()

    def _frontend(self, batch, last_state, last_delta):
        x0, factors, viz_out = self.visual_frontend(batch)
        self.last_state = x0
        return False if x0 is None else (x0, factors, viz_out)

    def _backend(self, factor_graph, x0):
        return x0, None          # iSAM2 over an empty graph: nothing to solve


class SlamModule(_WorkerModule):
    attr = "slam"

    def build(self):
        if self.name != "VioSLAM":
            raise NotImplementedError(self.name)
        return VioSLAM(self.name, self.args, self.device)

    def spin_once(self, inp):
        out = self.slam(inp)
        if not out or self.slam.stop_condition():
            self.shutdown_module()
        return out


class FusionModule(_WorkerModule):
    attr = "fusion"

    def build(self):
        if self.name in ("tsdf", "sigma"):
            from .tsdf_fusion import TsdfFusion
            return TsdfFusion(self.name, self.args, self.device)
        if self.name == "nerf":
            from .nerf_fusion import NerfFusion
            return NerfFusion(self.name, self.args, self.device)
        raise NotImplementedError(self.name)

    def get_input_packet(self, timeout=1e-10):
        got = super().get_input_packet(timeout=timeout)          # never block the trainer on an empty queue
        return got if got is not None else False                 # `False` = "no packet: keep fitting" (fusion_module.py:31-33)

    def spin_once(self, packet):
        out = self.fusion.fuse(packet)
        if self.fusion.stop_condition():
            print("Stopping fusion module!")
            self.shutdown_module()
        return out
