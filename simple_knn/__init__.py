"""Import shim: ``from simple_knn._C import distCUDA2`` resolves to the sm_100a implementation in
gaussianip_b200 (replaces the reference's vendored extension gaussiansplatting/submodules/simple-knn)."""
