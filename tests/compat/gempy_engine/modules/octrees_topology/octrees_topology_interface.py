def get_regular_grid_value_for_level(octree_list, level=None, value_type=None, scalar_n=-1):
    raise NotImplementedError("legacy gp2/gp3 helper: read Solutions.raw_arrays instead")
