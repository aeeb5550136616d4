measure = None  # import-only stub
