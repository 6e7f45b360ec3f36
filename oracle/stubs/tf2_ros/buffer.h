// Minimal stand-in header so the reference sources compile without ROS 2 (oracle/_ref only).
#pragma once
namespace tf2_ros { class Buffer {}; }
