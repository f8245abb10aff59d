// solver.cu placeholder
