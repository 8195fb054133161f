// placeholder until the odometry driver restatement lands
