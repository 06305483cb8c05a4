from .cdl import (CDL, CDLRealization, CDLType, CdlChannelState, ClusterDelayLineSample, cdl_propagate_batch)

__all__ = ["CDL", "CDLRealization", "CDLType", "CdlChannelState", "ClusterDelayLineSample", "cdl_propagate_batch"]
