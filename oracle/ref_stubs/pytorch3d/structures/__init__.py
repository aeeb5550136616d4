class Meshes:  # import-only stub
    pass
