class Trimesh:  # import-only stub
    pass
