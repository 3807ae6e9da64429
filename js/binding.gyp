{
  "targets": [{
    "target_name": "gsplat_b200",
    "sources": ["gsplat_b200_addon.cc"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../gaussiansplats3d_b200/csrc", "-lgsplat_b200"],
    "ldflags": ["-Wl,-rpath,<(module_root_dir)/../gaussiansplats3d_b200/csrc"],
    "cflags_cc": ["-std=c++17"]
  }]
}
