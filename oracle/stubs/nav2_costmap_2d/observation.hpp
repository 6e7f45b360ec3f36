// Minimal stand-in header so the reference sources compile without ROS 2 (oracle/_ref only).
// TEST INFRASTRUCTURE: declares just the names /root/reference/src/sfw_planner.cpp and its headers use.
#pragma once
