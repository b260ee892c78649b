"""corenet_b200: B200-native (sm_100a) implementation of the CoReNet hot path.

Host-side mirror of the reference's module surface:
  corenet_b200.model.core_net.CoreNet          <- corenet.model.core_net.CoreNet
  corenet_b200.model.{resnet50,reconstruction_decoder,ray_traced_skip_connection,batch_renorm,losses}
  corenet_b200.cc.fill_voxels                  <- corenet.cc.fill_voxels
  corenet_b200.geometry.voxelization           <- corenet.geometry.voxelization
over the C-ABI library declared in include/corenet_b200.h.
"""
__version__ = "0.1.0"
