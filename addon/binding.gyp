{
  "targets": [{
    "target_name": "gravitas_b200",
    "sources": ["binding.cc"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../blackhole-simulation_b200/gravitas_b200", "-lgravitas_b200",
                  "-Wl,-rpath,<(module_root_dir)/../blackhole-simulation_b200/gravitas_b200"],
    "cflags_cc": ["-std=c++17"]
  }]
}
