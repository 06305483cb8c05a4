"""3GPP cluster-delay-line channel (static CDL-A..E) behind the reference's plugin API, computed on the B200.

Host-side mirror of hermespy/channel/cdl/cdl.py (``CDL`` :353-471, ``CDLRealization`` :174-303) and of the
sample class hermespy/channel/cdl/cluster_delay_lines.py:255-756 (``ClusterDelayLineSample``).  Realization and
sampling -- table look-ups, ray coupling permutations, Jones phases -- stay on the host with numpy (RNG parity);
the per-ray array responses and the propagation (cluster_delay_lines.py:409-558; 95 % of the reference's time,
SURVEY 3.4) go through ``hb_cdl_propagate_host`` / ``hb_cdl_state``.
"""
from __future__ import annotations

from enum import Enum
from math import ceil
from typing import List, Optional, Sequence, Set

import numpy as np

from ... import config
from ...core import InterpolationMode, SignalBlock
from ..channel import Channel, ChannelRealization, ChannelSample, ChannelSampleHook, LinkState
from ..consistent import ConsistentGenerator, ConsistentRealization, ConsistentUniform
from .tables import CLUSTERS, PER_CLUSTER, RAY_OFFSETS

#: Ray partitions of the two strongest clusters; indexed by *cluster* index in the reference
#: (cluster_delay_lines.py:258-262, 450-457 -- SURVEY F9), which this mirror reproduces.
SUBCLUSTER_INDICES: List[List[int]] = [[0, 1, 2, 3, 4, 5, 6, 7, 18, 19], [8, 9, 10, 11, 16, 17], [12, 13, 14, 15]]


class CDLType(Enum):
    A = 0
    B = 1
    C = 2
    D = 3
    E = 4


def _pose12(pose) -> np.ndarray:
    m = np.asarray(pose.matrix, dtype=np.float64)
    return np.concatenate([m[:3, :3].ravel(), m[:3, 3]])


class CdlChannelState(object):
    """Impulse-response channel state of a CDL sample per delay group.

    ``gains[G, Nrx, Ntx, T]`` (GPU result) and ``delays[G]``; ``dense_state()`` scatters them into the reference's
    dense ``[Nrx, Ntx, T, 1 + D]`` layout (cluster_delay_lines.py:568-590).
    """

    def __init__(self, gains: np.ndarray, delays: np.ndarray, num_taps: int) -> None:
        self.gains = gains
        self.delays = delays
        self.num_delay_taps = int(num_taps)

    def dense_state(self) -> np.ndarray:
        G, nrx, ntx, T = self.gains.shape
        raw = np.zeros((nrx, ntx, T, self.num_delay_taps), dtype=np.complex128)
        for g, d in enumerate(self.delays):
            raw[:, :, :, d] = self.gains[g]
        return raw


class ClusterDelayLineSample(ChannelSample):
    """Sample of a 3GPP cluster delay line channel (cluster_delay_lines.py:255-756); same constructor signature."""

    subcluster_indices = SUBCLUSTER_INDICES

    def __init__(self, line_of_sight: bool, rice_factor: float, azimuth_of_arrival: np.ndarray,
                 zenith_of_arrival: np.ndarray, azimuth_of_departure: np.ndarray, zenith_of_departure: np.ndarray,
                 delay_offset: float, cluster_delays: np.ndarray, cluster_delay_spread: float,
                 cluster_powers: np.ndarray, polarization_transformations: np.ndarray, state: LinkState) -> None:
        ChannelSample.__init__(self, state)
        self.__line_of_sight = line_of_sight
        self.__rice_factor = rice_factor
        self.__aoa = azimuth_of_arrival
        self.__zoa = zenith_of_arrival
        self.__aod = azimuth_of_departure
        self.__zod = zenith_of_departure
        self.__delay_offset = delay_offset
        self.__cluster_delays = cluster_delays
        self.__cluster_delay_spread = cluster_delay_spread
        self.__cluster_powers = cluster_powers
        self.__jones = polarization_transformations
        self.__num_clusters = cluster_delays.shape[0]
        self.__num_rays = azimuth_of_arrival.shape[1]
        self.__max_delay = (max(np.max(cluster_delays[:3] + cluster_delay_spread * 2.56), cluster_delays.max())
                            + self.__delay_offset)  # cluster_delay_lines.py:316-319
        self.__block = None

    line_of_sight = property(lambda self: self.__line_of_sight)
    rice_factor = property(lambda self: self.__rice_factor)
    azimuth_of_arrival = property(lambda self: self.__aoa)
    zenith_of_arrival = property(lambda self: self.__zoa)
    azimuth_of_departure = property(lambda self: self.__aod)
    zenith_of_departure = property(lambda self: self.__zod)
    cluster_delays = property(lambda self: self.__cluster_delays)
    cluster_delay_spread = property(lambda self: self.__cluster_delay_spread)
    cluster_powers = property(lambda self: self.__cluster_powers)
    polarization_transformations = property(lambda self: self.__jones)
    num_clusters = property(lambda self: self.__num_clusters)
    num_rays = property(lambda self: self.__num_rays)
    max_delay = property(lambda self: self.__max_delay)
    delay_offset = property(lambda self: self.__delay_offset)

    @property
    def expected_energy_scale(self) -> float:
        return float(np.sum(self.cluster_powers))

    @property
    def max_delay_in_samples(self) -> int:
        return ceil(self.max_delay * self.bandwidth)

    # ---- kernel parameter block ------------------------------------------------------------------------
    def ray_term_index(self):
        """(cluster index, ray index, delay seconds) of every term in the reference generator's order
        (cluster_delay_lines.py:425-457)."""
        C, R = self.__num_clusters, self.__num_rays
        nsplit = min(2, C)
        nvirtual = 3 * nsplit + max(0, C - 2)
        sub = (np.repeat(self.__cluster_delays[:nsplit, None], 3, axis=1)
               + self.__cluster_delay_spread * np.array([0.0, 1.28, 2.56]))
        vdelays = np.concatenate((sub.flatten(), self.__cluster_delays[nsplit:]))
        cs, rs, ds = [], [], []
        for v in range(nvirtual):
            c = int(v / 3) if v < 6 else v - 4
            rays = SUBCLUSTER_INDICES[c] if c < nsplit else range(R)
            for r in rays:
                cs.append(c)
                rs.append(r)
                ds.append(vdelays[v])
        return np.array(cs, dtype=np.int64), np.array(rs, dtype=np.int64), np.array(ds, dtype=np.float64)

    def kernel_block(self):
        """Flat parameter block of this sample for the CUDA kernels (``kernels.CdlBlock`` with B = 1)."""
        if self.__block is None:
            from ...kernels import CdlBlock

            fs = self.bandwidth
            c, r, d = self.ray_term_index()
            rice_lin = 10.0 ** (self.__rice_factor / 10.0)  # tools/math.py db2lin
            nlos_scale = (1.0 + rice_lin) ** -0.5 if self.__line_of_sight else 1.0
            term_delay = np.array([int((dd + self.__delay_offset) * fs) for dd in d], dtype=np.int32)  # truncation (:548)
            angles = np.stack([self.__aoa[c, r], self.__zoa[c, r], self.__aod[c, r], self.__zod[c, r]], axis=1)
            jones = np.ascontiguousarray(self.__jones[:, :, c, r].transpose(2, 0, 1))
            amp = np.sqrt(self.__cluster_powers[c] / self.__num_rays) * nlos_scale
            self.__block = CdlBlock(
                term_delay=term_delay, max_delay=self.max_delay_in_samples, angles=angles[None], jones=jones[None],
                amplitude=amp[None], tx_pose=_pose12(self.transmitter_pose)[None],
                rx_pose=_pose12(self.receiver_pose)[None],
                rel_velocity=(np.asarray(self.receiver_velocity, float) - np.asarray(self.transmitter_velocity, float))[None],
                tx_topology=self.transmitter_antennas.topology, rx_topology=self.receiver_antennas.topology,
                carrier_frequency=self.carrier_frequency, sampling_rate=fs, line_of_sight=bool(self.__line_of_sight),
                los_delay=int((self.__cluster_delays[0] + self.__delay_offset) * fs),
                los_amplitude=float((rice_lin / (1 + rice_lin)) ** 0.5))
        return self.__block

    # ---- plugin interface -------------------------------------------------------------------------------
    def _propagate(self, signal: SignalBlock, interpolation: InterpolationMode) -> SignalBlock:
        """GPU replacement of cluster_delay_lines.py:526-558 for one block."""
        from ...kernels import cdl_propagate_host

        T = signal.shape[1]
        D = self.max_delay_in_samples
        if interpolation != InterpolationMode.NEAREST:
            # the reference silently accumulates nothing for other modes (cluster_delay_lines.py:547) -- SURVEY F3
            out = np.zeros((self.num_receive_antennas, T + D), dtype=np.complex128)
        else:
            x = np.ascontiguousarray(np.asarray(signal, dtype=np.complex128))[None]
            out = cdl_propagate_host(x, self.kernel_block(), precision=config.precision)[0]
        return SignalBlock(out.shape[0], out.shape[1], getattr(signal, "offset", 0), out.tobytes())

    def state(self, num_samples: int, max_num_taps: int,
              interpolation_mode: InterpolationMode = InterpolationMode.NEAREST) -> CdlChannelState:
        """GPU replacement of cluster_delay_lines.py:561-592."""
        from ...kernels import CdlDeviceBlock, cdl_state

        D = min(max_num_taps, self.max_delay_in_samples)
        h, gd = cdl_state(CdlDeviceBlock(self.kernel_block(), device=f"cuda:{config.device}"), num_samples)
        h = h[0].cpu().numpy()
        keep = gd < max_num_taps  # cluster_delay_lines.py:583-584
        return CdlChannelState(h[keep], gd[keep], 1 + D)

    def reciprocal(self, state: LinkState) -> "ClusterDelayLineSample":
        """Arrival and departure angles swap (cluster_delay_lines.py:732-756)."""
        return ClusterDelayLineSample(self.line_of_sight, self.rice_factor, self.azimuth_of_departure,
                                      self.zenith_of_departure, self.azimuth_of_arrival, self.zenith_of_arrival,
                                      self.delay_offset, self.cluster_delays, self.cluster_delay_spread,
                                      self.cluster_powers, self.polarization_transformations, state)


def cdl_propagate_batch(samples: Sequence[ClusterDelayLineSample], signals: Sequence[np.ndarray],
                        precision: Optional[str] = None) -> List[np.ndarray]:
    """Propagate many (sample, signal) pairs with one pipeline call per array geometry and block length; samples with
    different delay structures (models, delay spreads, line-of-sight states) share it through per-link delay tables."""
    from ...kernels import CdlBlock, cdl_propagate_host

    precision = config.precision if precision is None else precision
    groups = {}
    for i, (s, x) in enumerate(zip(samples, signals)):
        x = np.asarray(x)
        key = (s.kernel_block().geometry_key(), x.shape, x.dtype.str)
        groups.setdefault(key, []).append(i)
    out: List[Optional[np.ndarray]] = [None] * len(samples)
    for idx in groups.values():
        blk = CdlBlock.stack([samples[i].kernel_block() for i in idx])
        x = np.stack([np.asarray(signals[i]) for i in idx])
        if x.dtype not in (np.complex64, np.complex128):
            x = x.astype(np.complex128)
        y = cdl_propagate_host(x, blk, precision=precision)
        for k, i in enumerate(idx):
            out[i] = y[k] if blk.link_max_delay is None else np.ascontiguousarray(y[k][:, : x.shape[2] + int(blk.link_max_delay[k])])
    return out  # type: ignore[return-value]


class CDLRealization(ChannelRealization[ClusterDelayLineSample]):
    """Realization of a static CDL model (cdl.py:174-303)."""

    def __init__(self, type: CDLType, rms_delay: float, rayleigh_factor: float, angle_coupling_indices: np.ndarray,
                 consistent_realization: ConsistentRealization, xpr_phase: ConsistentUniform,
                 sample_hooks: Set[ChannelSampleHook], gain: float) -> None:
        ChannelRealization.__init__(self, sample_hooks, gain)
        self.__type = type
        self.__rms_delay = rms_delay
        self.__rayleigh_factor = rayleigh_factor
        self.__coupling = angle_coupling_indices
        self.__consistent_realization = consistent_realization
        self.__xpr_phase = xpr_phase

    def _sample(self, state: LinkState) -> ClusterDelayLineSample:
        cs = self.__consistent_realization.sample(state.transmitter.position, state.receiver.position)
        tab = CLUSTERS[self.__type.value]
        c_asd, c_asa, c_zsd, c_zsa, xpr_db, los = PER_CLUSTER[self.__type.value, :]
        cluster_powers = 10 ** (tab[:, 1] / 10)
        cluster_delays = self.__rms_delay * tab[:, 0]
        # equation 7.7-0a: ray angles = cluster angle + spread * offset; 7.7-0b: random coupling per cluster
        aod = np.take_along_axis(np.add.outer(tab[:, 2], c_asd * RAY_OFFSETS), self.__coupling[0, :], axis=1)
        aoa = np.take_along_axis(np.add.outer(tab[:, 3], c_asa * RAY_OFFSETS), self.__coupling[1, :], axis=1)
        zod = np.take_along_axis(np.add.outer(tab[:, 4], c_zsd * RAY_OFFSETS), self.__coupling[2, :], axis=1)
        zoa = np.take_along_axis(np.add.outer(tab[:, 5], c_zsa * RAY_OFFSETS), self.__coupling[3, :], axis=1)
        xpf = 10 ** (xpr_db / 10)
        jones = np.exp(2j * np.pi * self.__xpr_phase.sample(cs))
        jones[0, 1, ::] *= xpf
        jones[1, 0, ::] *= xpf
        return ClusterDelayLineSample(bool(los), self.__rayleigh_factor, np.pi / 180 * aoa, np.pi / 180 * zoa,
                                      np.pi / 180 * aod, np.pi / 180 * zod, 0, cluster_delays, self.__rms_delay,
                                      cluster_powers, jones, state)

    def _reciprocal_sample(self, sample: ClusterDelayLineSample, state: LinkState) -> ClusterDelayLineSample:
        return sample.reciprocal(state)


class CDL(Channel[CDLRealization, ClusterDelayLineSample]):
    """Static cluster delay line model for link-level simulations (cdl.py:353-471)."""

    def __init__(self, model_type: CDLType, rms_delay: float, rayleigh_factor: float = 0.0,
                 decorrelation_distance: float = 30.0, **kwargs) -> None:
        Channel.__init__(self, **kwargs)
        self.__model_type = CDLType(model_type) if not isinstance(model_type, CDLType) else model_type
        self.rms_delay = rms_delay
        self.rayleigh_factor = rayleigh_factor
        self.decorrelation_distance = decorrelation_distance
        self.__generator = ConsistentGenerator(self)
        self.__xpr_phase = self.__generator.uniform((2, 2, CLUSTERS[self.__model_type.value].shape[0], RAY_OFFSETS.size))

    model_type = property(lambda self: self.__model_type)

    @property
    def rms_delay(self) -> float:
        return self.__rms_delay

    @rms_delay.setter
    def rms_delay(self, value: float) -> None:
        if value < 0:
            raise ValueError("The delay spread must be non-negative.")
        self.__rms_delay = value

    @property
    def rayleigh_factor(self) -> float:
        return self.__rayleigh_factor

    @rayleigh_factor.setter
    def rayleigh_factor(self, value: float) -> None:
        if value < 0:
            raise ValueError("The K-factor must be non-negative.")
        self.__rayleigh_factor = value

    @property
    def decorrelation_distance(self) -> float:
        return self.__decorrelation_distance

    @decorrelation_distance.setter
    def decorrelation_distance(self, value: float) -> None:
        if value < 0:
            raise ValueError("The decorrelation distance must be non-negative.")
        self.__decorrelation_distance = value

    def _realize(self) -> CDLRealization:
        # 4 x C permutations drawn in this order (cdl.py:449-460), then the consistent realization
        candidates = np.arange(RAY_OFFSETS.size)
        num_clusters = CLUSTERS[self.__model_type.value].shape[0]
        coupling = np.array([[self._rng.permutation(candidates) for _ in range(num_clusters)] for _ in range(4)])
        return CDLRealization(self.__model_type, self.rms_delay, self.rayleigh_factor, coupling,
                              self.__generator.realize(self.decorrelation_distance), self.__xpr_phase,
                              self.sample_hooks, self.gain)
